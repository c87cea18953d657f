"""Flax-msgpack checkpoints of the flat device state, so ``--resume`` and the evaluators of the reference can read
weights trained here and this build can continue from the reference's ``step_XXXXX.msgpack`` files.

ref: save_checkpoint / load_checkpoint trainers/utils.py:45-58, unwrapped (params only) form :548-586,
get_checkpoint_step :535-545, call sites trainers/train.py:266-276,406-408,464-466.

Wire format (flax 0.10.4 ``flax.serialization``, pinned in uv.lock:327-328; restated here because flax is not
installable in this image -- parity unpinned):

* ``to_bytes(x) = msgpack.packb(to_state_dict(x), default=ext_pack, strict_types=True)``;
* an array leaf is ``ExtType(1, msgpack.packb((shape, dtype.name, C-order bytes), use_bin_type=True))``, a NumPy
  scalar is ``ExtType(3, <same triple>)``, Python ints / floats travel as native msgpack numbers;
* a ``TrainState`` is the dict of its pytree fields ``{"step", "params", "opt_state"}`` (``apply_fn`` and ``tx`` are
  static); tuples become ``{"0": ..., "1": ...}`` and NamedTuples become dicts of their fields, so the state of
  ``optax.adamw`` = ``chain(scale_by_adam, add_decayed_weights, scale_by_learning_rate)`` (optax 0.2.5) is
  ``{"0": {"count": int32[], "mu": <params tree>, "nu": <params tree>}, "1": {}, "2": {}}``.

Host-side format code only: the arrays are copied device <-> host once per save / load, nothing here is on the
hot path.
"""
from __future__ import annotations

import re
from pathlib import Path

import msgpack
import numpy as np
import torch

from .mlp_flow import FlatParams, TrainState

_EXT_NDARRAY, _EXT_NPSCALAR = 1, 3
_MAX_CHUNK_BYTES = 2 ** 30   # flax splits larger leaves into "__msgpack_chunked_array__" dicts


def _pack_array(a: np.ndarray) -> bytes:
    return msgpack.packb((list(a.shape), a.dtype.name, a.tobytes("C")), use_bin_type=True)


def _ext_pack(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    if isinstance(x, np.ndarray):
        if x.nbytes > _MAX_CHUNK_BYTES:
            raise ValueError("leaves above 1 GiB need flax's chunked encoding, which this writer does not emit")
        return msgpack.ExtType(_EXT_NDARRAY, _pack_array(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _pack_array(np.asarray(x)))
    raise TypeError(f"cannot serialise {type(x)}")


def _ext_unpack(code, data):
    if code in (_EXT_NDARRAY, _EXT_NPSCALAR):
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        a = np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape)
        return a[()] if code == _EXT_NPSCALAR else a
    return msgpack.ExtType(code, data)


def _unchunk(node):
    """Re-assemble flax's chunked big-array leaves (reader side only)."""
    if isinstance(node, dict):
        if node.get("__msgpack_chunked_array__"):
            shape = [node["shape"][str(i)] for i in range(len(node["shape"]))]
            chunks = [node["chunks"][str(i)] for i in range(len(node["chunks"]))]
            return np.concatenate([np.asarray(c).ravel() for c in chunks]).reshape(shape)
        return {k: _unchunk(v) for k, v in node.items()}
    return node


def msgpack_serialize(tree) -> bytes:
    return msgpack.packb(tree, default=_ext_pack, strict_types=True)


def msgpack_restore(data: bytes):
    return _unchunk(msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


# ------------------------------------------------------------------------------------------------ state dicts
def _tree_to_numpy(model, flat: torch.Tensor) -> dict:
    host = flat.detach().to("cpu", torch.float32).numpy()
    root: dict = {}
    for path, (off, shape) in model.leaf_slices().items():
        d = root
        for p in path[:-1]:
            d = d.setdefault(p, {})
        d[path[-1]] = host[off:off + int(np.prod(shape))].reshape(shape)
    return root


def _numpy_to_flat(model, tree: dict, what: str) -> torch.Tensor:
    """Flax ``from_state_dict`` semantics: the stored tree must have exactly the template's structure."""
    out = np.empty(model.param_count(), dtype=np.float32)
    seen = 0
    for path, (off, shape) in model.leaf_slices().items():
        d = tree
        for p in path:
            if not isinstance(d, dict) or p not in d:
                raise ValueError(f"{what}: missing leaf {'/'.join(path)} in the checkpoint")
            d = d[p]
        a = np.asarray(d)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{what}: leaf {'/'.join(path)} has shape {tuple(a.shape)}, the model expects {tuple(shape)}")
        out[off:off + a.size] = a.astype(np.float32, copy=False).ravel()
        seen += 1

    def count(n):
        return sum(count(v) for v in n.values()) if isinstance(n, dict) else 1
    if count(tree) != seen:
        raise ValueError(f"{what}: the checkpoint holds {count(tree)} leaves, the model has {seen}")
    return torch.from_numpy(out)


def to_state_dict(state: TrainState) -> dict:
    """``flax.serialization.to_state_dict(TrainState)`` for the AdamW state this build keeps flat on the device."""
    model = state.model
    fp = model.flat_params(state.params)
    return {
        "step": int(state.step),
        "params": _tree_to_numpy(model, fp.flat),
        "opt_state": {
            "0": {"count": np.asarray(int(state.opt_state["count"]), dtype=np.int32),
                  "mu": _tree_to_numpy(model, state.opt_state["mu"]),
                  "nu": _tree_to_numpy(model, state.opt_state["nu"])},
            "1": {},
            "2": {},
        },
    }


def from_state_dict(state_template: TrainState, sd: dict) -> TrainState:
    model = state_template.model
    for k in ("step", "params", "opt_state"):
        if k not in sd:
            raise ValueError(f"checkpoint has no '{k}' entry (keys: {sorted(sd)})")
    dev = model.flat_params(state_template.params).flat.device
    adam = sd["opt_state"].get("0") if isinstance(sd["opt_state"], dict) else None
    if not isinstance(adam, dict) or not {"count", "mu", "nu"} <= set(adam):
        raise ValueError("checkpoint opt_state is not optax.adamw's (ScaleByAdamState, EmptyState, EmptyState)")
    flat = _numpy_to_flat(model, sd["params"], "params").to(dev)
    opt_state = {"count": int(np.asarray(adam["count"])),
                 "mu": _numpy_to_flat(model, adam["mu"], "opt_state.mu").to(dev),
                 "nu": _numpy_to_flat(model, adam["nu"], "opt_state.nu").to(dev)}
    return TrainState(int(np.asarray(sd["step"])), state_template.apply_fn, FlatParams(model, flat).tree(),
                      state_template.tx, opt_state, model)


def to_bytes(target) -> bytes:
    """TrainState -> bytes, or a params tree (``unwrap_checkpoint`` form) -> bytes."""
    if isinstance(target, TrainState):
        return msgpack_serialize(to_state_dict(target))
    return msgpack_serialize(_to_numpy_tree(target))


def from_bytes(target, data: bytes):
    """``flax.serialization.from_bytes``: restores into the structure of ``target`` (a TrainState template), or
    returns the raw restored dict when ``target`` is an empty dict (trainers/utils.py:575-586)."""
    sd = msgpack_restore(data)
    if isinstance(target, TrainState):
        return from_state_dict(target, sd)
    return sd


def _to_numpy_tree(node):
    if isinstance(node, dict):
        return {k: _to_numpy_tree(v) for k, v in node.items()}
    if isinstance(node, torch.Tensor):
        return node.detach().cpu().numpy()
    return node


# ------------------------------------------------------------------------------------------------ files
def save_checkpoint(path: Path, state: TrainState) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with path.open("wb") as f:
        f.write(to_bytes(state))


def load_checkpoint(path: Path, state_template: TrainState) -> TrainState:
    with Path(path).open("rb") as f:
        return from_bytes(state_template, f.read())


def unwrap_checkpoint(state: TrainState) -> dict:
    return {"params": state.params}


def save_unwrapped_checkpoint(path: Path, params: dict) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with path.open("wb") as f:
        f.write(to_bytes(params))


def load_unwrapped_checkpoint(path: Path) -> dict:
    with Path(path).open("rb") as f:
        return from_bytes({}, f.read())


def get_checkpoint_step(checkpoint_path: Path) -> int:
    match = re.search(r"step_(\d+)\.msgpack", Path(checkpoint_path).name)
    if match:
        return int(match.group(1))
    raise ValueError(f"Could not extract step number from: {checkpoint_path}")

"""Host-side mirror of ``models/mlp_flow.py`` (ConditionalFlow) and ``models/train_state.py`` on libmfac.

    model = ConditionalFlow(noise_dimension, condition_dimension, num_blocks, latent_dimension)
    params = model.init(seed)["params"]                       # Flax-shaped tree of CUDA tensors
    latents = model.apply({"params": params}, x, method="encode")
    out = model.apply({"params": params}, x, time, latents)   # latents=None -> zeros

The tree's leaves are views into ONE flat fp32 buffer in jax tree_flatten order (the layout the
C ABI consumes); any other tree with the same structure is accepted and flattened by copy.
ref: models/mlp_flow.py:125-230, models/train_state.py:4, trainers/train.py:229-264.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib

_BLOCK_LEAVES = [
    ("conditioning_layer", "dense1", "bias"), ("conditioning_layer", "dense1", "kernel"),
    ("conditioning_layer", "dense2", "bias"), ("conditioning_layer", "dense2", "kernel"),
    ("mlp", "dense1", "bias"), ("mlp", "dense1", "kernel"),
    ("mlp", "dense2", "bias"), ("mlp", "dense2", "kernel"),
]
_ENC_LEAVES = [("encoder_mlp", "dense1", "bias"), ("encoder_mlp", "dense1", "kernel"),
               ("encoder_mlp", "dense2", "bias"), ("encoder_mlp", "dense2", "kernel")]


class ParamTree(dict):
    """Nested dict of tensors that remembers the flat buffer its leaves alias."""
    flat: torch.Tensor | None = None
    owner: "FlatParams | None" = None


class FlatParams:
    """Flat fp32 parameter vector + its bf16 shadow (kernels) for one ConditionalFlow geometry."""

    def __init__(self, model: "ConditionalFlow", flat: torch.Tensor):
        self.model = model
        self.flat = flat
        self._shadow = None
        self._shadow_version = None

    def leaf_slices(self):
        return self.model.leaf_slices()

    def tree(self) -> ParamTree:
        root = ParamTree()
        for path, (off, shape) in self.leaf_slices().items():
            d = root
            for p in path[:-1]:
                d = d.setdefault(p, {})
            n = 1
            for s in shape:
                n *= s
            d[path[-1]] = self.flat[off:off + n].view(shape)
        root.flat = self.flat
        root.owner = self
        return root

    def shadow(self) -> torch.Tensor:
        """bf16 kernels + padded biases for the tensor-core GEMMs, re-cast when the fp32 buffer changed."""
        ver = self.flat._version
        if self._shadow is None or self._shadow_version != ver:
            l = _lib.lib()
            dims = self.model.dims
            if self._shadow is None:
                nbytes = l.mfac_mlp_shadow_bytes(C.byref(dims))
                self._shadow = torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
            with torch.cuda.device(self.flat.device):
                _lib.check(l.mfac_mlp_cast_params(C.byref(dims), self.flat.data_ptr(), self._shadow.data_ptr(),
                                                  _lib.stream_ptr()), "cast_params")
            self._shadow_version = ver
        return self._shadow

    def mark_shadow_current(self):
        """Called after mfac_adamw_step refreshed the shadow in the same pass."""
        self._shadow_version = self.flat._version


class ConditionalFlow:
    def __init__(self, noise_dimension: int, condition_dimension: int, num_blocks: int, latent_dimension: int):
        if condition_dimension % 2:
            raise ValueError(f"condition_dimension must be even, got {condition_dimension}")
        self.noise_dimension = int(noise_dimension)
        self.condition_dimension = int(condition_dimension)
        self.num_blocks = int(num_blocks)
        self.latent_dimension = int(latent_dimension)
        self.dims = _lib.MlpDims(self.noise_dimension, self.latent_dimension, self.condition_dimension, self.num_blocks)
        self._slices = None
        self._ws = {}

    # ---------------------------------------------------------------- layout
    def param_count(self) -> int:
        D, L, Cd, nb = self.noise_dimension, self.latent_dimension, self.condition_dimension, self.num_blocks
        I, He = L + D, (D + L) // 2
        blk = Cd + Cd * Cd + (2 * I + D) + Cd * (2 * I + D) + I + I * I + D + I * D
        return nb * blk + He + D * He + L + He * L

    def leaf_slices(self) -> dict:
        """path tuple -> (offset, shape); same order as include/mfac.h documents."""
        if self._slices is None:
            D, L, Cd, nb = self.noise_dimension, self.latent_dimension, self.condition_dimension, self.num_blocks
            I, He = L + D, (D + L) // 2
            shapes_blk = [(Cd,), (Cd, Cd), (2 * I + D,), (Cd, 2 * I + D), (I,), (I, I), (D,), (I, D)]
            shapes_enc = [(He,), (D, He), (L,), (He, L)]
            out, off = {}, 0
            for k in range(nb):
                for leaf, shp in zip(_BLOCK_LEAVES, shapes_blk):
                    out[(f"blocks_{k}",) + leaf] = (off, shp)
                    off += math.prod(shp)
            for leaf, shp in zip(_ENC_LEAVES, shapes_enc):
                out[("encoder",) + leaf] = (off, shp)
                off += math.prod(shp)
            self._slices = out
        return self._slices

    # ---------------------------------------------------------------- params
    def init(self, key=0, *args, device="cuda", on_device: bool = False, **kwargs) -> dict:
        """Flax-style init: lecun_normal kernels (truncated normal, var 1/fan_in), zero biases.
        ``key`` is an int seed or a torch.Generator (the reference passes a jax PRNGKey).
        ``on_device=True`` draws the kernels with the device's generator instead of the host's (same distribution, another
        stream): a 1 G-parameter geometry initialises in milliseconds instead of a minute of host RNG + upload."""
        if on_device:
            gen = torch.Generator(device=device).manual_seed(int(key) if not isinstance(key, torch.Generator) else key.initial_seed())
            flat = torch.zeros(self.param_count(), dtype=torch.float32, device=device)
            for path, (off, shp) in self.leaf_slices().items():
                if path[-1] == "kernel":
                    w = flat[off:off + math.prod(shp)]
                    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
                    w.mul_(math.sqrt(1.0 / shp[0]) / 0.87962566103423978)
            return {"params": FlatParams(self, flat).tree()}
        gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
        flat = torch.zeros(self.param_count(), dtype=torch.float32)
        for path, (off, shp) in self.leaf_slices().items():
            if path[-1] == "kernel":
                w = torch.empty(shp, dtype=torch.float32)
                torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
                flat[off:off + w.numel()] = (w * (math.sqrt(1.0 / shp[0]) / 0.87962566103423978)).reshape(-1)
        return {"params": FlatParams(self, flat.to(device)).tree()}

    def flat_params(self, params) -> FlatParams:
        """Accepts a ParamTree (zero-copy), a FlatParams, a flat tensor, or any Flax-shaped tree of tensors."""
        if isinstance(params, FlatParams):
            return params
        if isinstance(params, ParamTree) and params.owner is not None:
            return params.owner
        if isinstance(params, torch.Tensor):
            return FlatParams(self, _lib.require_cuda(params, "params").to(torch.float32).contiguous())
        parts = []
        for path, (off, shp) in self.leaf_slices().items():
            d = params
            for p in path:
                d = d[p]
            if tuple(d.shape) != tuple(shp):
                raise ValueError(f"parameter {'/'.join(path)} has shape {tuple(d.shape)}, expected {shp}")
            parts.append(_lib.require_cuda(d, "/".join(path)).to(torch.float32).reshape(-1))
        return FlatParams(self, torch.cat(parts))

    def workspace(self, kind: int, B: int, device) -> torch.Tensor:
        key = (kind, int(B), str(device), _lib.layout_epoch)
        ws = self._ws.get(key)
        if ws is None:
            n = _lib.lib().mfac_workspace_bytes(kind, C.byref(self.dims), int(B))
            if n == 0:
                raise _lib.MfacError("mfac_workspace_bytes returned 0 (bad dimensions)")
            ws = torch.empty(n, dtype=torch.uint8, device=device)
            self._ws = {k: v for k, v in self._ws.items() if k[0] != kind}  # keep one per kind
            self._ws[key] = ws
        return ws

    # ---------------------------------------------------------------- apply
    def apply(self, variables, x, time=None, latents=None, method=None):
        fp = self.flat_params(variables["params"])
        if method == "encode":
            return self.encode(fp, x)
        if method is not None:
            raise ValueError(f"unknown method {method!r}")
        if time is None:
            raise TypeError("apply() missing required argument: 'time'")
        return self.decode(fp, x, time, latents)

    def encode(self, fp: FlatParams, x: torch.Tensor) -> torch.Tensor:
        x = _lib.require_cuda(x, "x").to(torch.float32).contiguous()
        if x.ndim != 2 or x.shape[1] != self.noise_dimension:
            raise ValueError(f"x must be [B, {self.noise_dimension}], got {tuple(x.shape)}")
        B = x.shape[0]
        out = torch.empty((B, self.latent_dimension), dtype=torch.float32, device=x.device)
        ws = self.workspace(_lib.WS_FORWARD, B, x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mfac_mlp_encode(C.byref(self.dims), fp.flat.data_ptr(), fp.shadow().data_ptr(),
                                                  x.data_ptr(), out.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                  _lib.stream_ptr()), "mlp_encode")
        return out

    def decode(self, fp: FlatParams, x, time, latents=None) -> torch.Tensor:
        x = _lib.require_cuda(x, "x").to(torch.float32).contiguous()
        time = _lib.require_cuda(time, "time").to(torch.float32).contiguous()
        B = x.shape[0]
        if x.ndim != 2 or x.shape[1] != self.noise_dimension:
            raise ValueError(f"x must be [B, {self.noise_dimension}], got {tuple(x.shape)}")
        if tuple(time.shape) != (B, 2):
            raise ValueError(f"time must be [B, 2] = (t, h), got {tuple(time.shape)}")
        lat_ptr = None
        if latents is not None:
            latents = _lib.require_cuda(latents, "latents").to(torch.float32).contiguous()
            if tuple(latents.shape) != (B, self.latent_dimension):
                raise ValueError(f"latents must be [B, {self.latent_dimension}], got {tuple(latents.shape)}")
            lat_ptr = latents.data_ptr()
        out = torch.empty_like(x)
        ws = self.workspace(_lib.WS_FORWARD, B, x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mfac_mlp_forward(C.byref(self.dims), fp.flat.data_ptr(), fp.shadow().data_ptr(),
                                                   x.data_ptr(), time.data_ptr(), lat_ptr, out.data_ptr(), B,
                                                   ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mlp_forward")
        return out

    def __call__(self, variables, x, time, latents=None):
        return self.apply(variables, x, time, latents)


class AdamW:
    """optax.adamw(learning_rate, weight_decay) hyper-parameters (trainers/train.py:236)."""

    def __init__(self, learning_rate: float = 1e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
                 weight_decay: float = 1e-4):
        self.learning_rate, self.b1, self.b2, self.eps, self.weight_decay = learning_rate, b1, b2, eps, weight_decay


def adamw(learning_rate: float, weight_decay: float = 1e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> AdamW:
    return AdamW(learning_rate, b1, b2, eps, weight_decay)


class TrainState:
    """flax.training.train_state.TrainState shape: step, apply_fn, params, tx, opt_state."""

    def __init__(self, step, apply_fn, params, tx, opt_state, model):
        self.step, self.apply_fn, self.params, self.tx, self.opt_state, self.model = step, apply_fn, params, tx, opt_state, model

    @classmethod
    def create(cls, *, apply_fn, params, tx: AdamW):
        model = getattr(apply_fn, "__self__", None)
        if not isinstance(model, ConditionalFlow):
            raise TypeError("apply_fn must be the bound ``apply`` of a ConditionalFlow")
        fp = model.flat_params(params)
        opt_state = {"count": 0, "mu": torch.zeros_like(fp.flat), "nu": torch.zeros_like(fp.flat)}
        return cls(0, apply_fn, fp.tree(), tx, opt_state, model)

    def apply_gradients(self, *, grads, grad_scale: float = 1.0, count_tensor=None, scratch=None):
        """AdamW update in place on the flat buffers (one fused kernel that also refreshes the bf16
        shadow); returns the advanced state like the reference's functional API.
        ``count_tensor`` (uint64 CUDA scalar) + ``scratch`` (2 fp32) select the graph-capturable form whose step
        count lives, and is advanced, on the device."""
        fp = self.model.flat_params(self.params)
        g = grads.flat if isinstance(grads, ParamTree) else self.model.flat_params(grads).flat
        tx = self.tx
        if count_tensor is not None:
            with torch.cuda.device(fp.flat.device):
                _lib.check(_lib.lib().mfac_adamw_step_dev(C.byref(self.model.dims), fp.flat.data_ptr(), g.data_ptr(),
                                                          self.opt_state["mu"].data_ptr(), self.opt_state["nu"].data_ptr(),
                                                          fp.shadow().data_ptr(), count_tensor.data_ptr(), scratch.data_ptr(),
                                                          tx.learning_rate, tx.b1, tx.b2, tx.eps, tx.weight_decay,
                                                          float(grad_scale), _lib.stream_ptr()), "adamw_step_dev")
            fp.mark_shadow_current()
            return self
        with torch.cuda.device(fp.flat.device):
            _lib.check(_lib.lib().mfac_adamw_step(C.byref(self.model.dims), fp.flat.data_ptr(), g.data_ptr(),
                                                  self.opt_state["mu"].data_ptr(), self.opt_state["nu"].data_ptr(),
                                                  fp.shadow().data_ptr(), int(self.opt_state["count"]),
                                                  tx.learning_rate, tx.b1, tx.b2, tx.eps, tx.weight_decay,
                                                  float(grad_scale), _lib.stream_ptr()), "adamw_step")
        fp.mark_shadow_current()
        self.opt_state["count"] += 1
        return TrainState(self.step + 1, self.apply_fn, self.params, self.tx, self.opt_state, self.model)

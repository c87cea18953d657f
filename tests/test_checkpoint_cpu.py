"""CPU tests of the Flax-msgpack checkpoint format (trainers/utils.py:45-58,548-586): byte layout against an independent
statement of flax.serialization's encoding, round trip of the flat device state, and from_bytes' structure errors."""
import struct

import msgpack
import numpy as np
import pytest
import torch

import meanflow_audio_codec_b200 as m
from meanflow_audio_codec_b200 import checkpoint as ck


def _state(D=8, L=16, C=8, nb=2, seed=0):
    model = m.ConditionalFlow(noise_dimension=D, condition_dimension=C, num_blocks=nb, latent_dimension=L)
    state = m.TrainState.create(apply_fn=model.apply, params=model.init(seed, device="cpu")["params"], tx=m.adamw(1e-4, 1e-4))
    g = torch.Generator().manual_seed(seed + 1)
    state.opt_state["mu"].copy_(torch.randn(model.param_count(), generator=g))
    state.opt_state["nu"].copy_(torch.rand(model.param_count(), generator=g))
    state.opt_state["count"] = 7
    state.step = 7
    return model, state


def _flax_leaf(a):
    """flax.serialization._ndarray_to_bytes + ExtType(1): msgpack ext 1 wrapping packb((shape, dtype.name, bytes))."""
    return msgpack.ExtType(1, msgpack.packb((a.shape, a.dtype.name, a.tobytes("C")), use_bin_type=True))


def test_bytes_are_flax_msgpack():
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    ours = ck.to_bytes({"params": {"dense": {"kernel": torch.from_numpy(a), "bias": torch.zeros(3)}}})
    theirs = msgpack.packb({"params": {"dense": {"kernel": _flax_leaf(a), "bias": _flax_leaf(np.zeros(3, np.float32))}}},
                           strict_types=True)
    assert ours == theirs
    # the ext payload really is (shape, dtype name, raw C-order bytes)
    ext = msgpack.unpackb(ours, raw=False)["params"]["dense"]["kernel"]
    assert ext.code == 1
    shape, name, buf = msgpack.unpackb(ext.data, raw=False)
    assert shape == [2, 3] and name == "float32" and buf == struct.pack("<6f", *range(6))
    back = ck.from_bytes({}, ours)
    np.testing.assert_array_equal(back["params"]["dense"]["kernel"], a)


def test_train_state_layout_is_flax_train_state_with_optax_adamw():
    model, state = _state()
    raw = msgpack.unpackb(ck.to_bytes(state), raw=False)
    assert list(raw) == ["step", "params", "opt_state"] and raw["step"] == 7
    assert sorted(raw["opt_state"]) == ["0", "1", "2"] and raw["opt_state"]["1"] == {} and raw["opt_state"]["2"] == {}
    adam = raw["opt_state"]["0"]
    assert sorted(adam) == ["count", "mu", "nu"]
    shape, name, buf = msgpack.unpackb(adam["count"].data, raw=False)
    assert (shape, name, struct.unpack("<i", buf)[0]) == ([], "int32", 7)     # ScaleByAdamState.count: int32 scalar array
    assert sorted(raw["params"]) == ["blocks_0", "blocks_1", "encoder"]
    assert sorted(raw["params"]["blocks_0"]) == ["conditioning_layer", "mlp"]
    assert sorted(raw["params"]["encoder"]["encoder_mlp"]["dense1"]) == ["bias", "kernel"]
    shape, name, _ = msgpack.unpackb(raw["params"]["blocks_1"]["mlp"]["dense2"]["kernel"].data, raw=False)
    assert shape == [16 + 8, 8] and name == "float32"                          # [in, out] as flax.linen.Dense stores it
    assert raw["params"].keys() == adam["mu"].keys() == adam["nu"].keys()


def test_round_trip_and_resume(tmp_path):
    model, state = _state()
    path = tmp_path / "checkpoints" / "step_00007.msgpack"
    ck.save_checkpoint(path, state)
    assert ck.get_checkpoint_step(path) == 7
    model2 = m.ConditionalFlow(8, 8, 2, 16)
    template = m.TrainState.create(apply_fn=model2.apply, params=model2.init(99, device="cpu")["params"], tx=m.adamw(1e-4, 1e-4))
    back = ck.load_checkpoint(path, template)
    assert back.step == 7 and back.opt_state["count"] == 7 and back.tx is template.tx
    assert torch.equal(model2.flat_params(back.params).flat, model.flat_params(state.params).flat)
    assert torch.equal(back.opt_state["mu"], state.opt_state["mu"]) and torch.equal(back.opt_state["nu"], state.opt_state["nu"])
    # params-only form (trainers/utils.py:548-586)
    ck.save_unwrapped_checkpoint(tmp_path / "params.msgpack", ck.unwrap_checkpoint(state))
    raw = ck.load_unwrapped_checkpoint(tmp_path / "params.msgpack")
    np.testing.assert_array_equal(raw["params"]["blocks_0"]["mlp"]["dense1"]["bias"],
                                  state.params["blocks_0"]["mlp"]["dense1"]["bias"].numpy())
    with pytest.raises(ValueError):
        ck.get_checkpoint_step(tmp_path / "params.msgpack")


def test_reads_a_reference_written_file():
    """A checkpoint as flax would write it (jnp int32 step scalar as ext 3, frozen-dict order, a chunked leaf)."""
    model, state = _state()
    sd = ck.to_state_dict(state)

    def enc(n):
        return {k: enc(v) for k, v in n.items()} if isinstance(n, dict) else _flax_leaf(np.asarray(n))
    tree = enc(sd)
    tree["step"] = msgpack.ExtType(3, msgpack.packb(((), "int32", np.int32(7).tobytes()), use_bin_type=True))
    k = sd["params"]["encoder"]["encoder_mlp"]["dense1"]["kernel"]
    halves = np.array_split(k.ravel(), 2)
    tree["params"]["encoder"]["encoder_mlp"]["dense1"]["kernel"] = {
        "__msgpack_chunked_array__": True, "shape": {str(i): s for i, s in enumerate(k.shape)},
        "chunks": {str(i): _flax_leaf(h) for i, h in enumerate(halves)}}
    tree["params"] = dict(reversed(list(tree["params"].items())))
    back = ck.from_bytes(state, msgpack.packb(tree, strict_types=True))
    assert back.step == 7
    assert torch.equal(model.flat_params(back.params).flat, model.flat_params(state.params).flat)


def test_structure_mismatch_raises_like_from_state_dict():
    model, state = _state()
    other = m.ConditionalFlow(8, 8, 3, 16)   # one more block
    template = m.TrainState.create(apply_fn=other.apply, params=other.init(0, device="cpu")["params"], tx=m.adamw(1e-4))
    with pytest.raises(ValueError, match="missing leaf"):
        ck.from_bytes(template, ck.to_bytes(state))
    wide = m.ConditionalFlow(16, 8, 2, 16)
    template = m.TrainState.create(apply_fn=wide.apply, params=wide.init(0, device="cpu")["params"], tx=m.adamw(1e-4))
    with pytest.raises(ValueError, match="shape"):
        ck.from_bytes(template, ck.to_bytes(state))
    small = m.ConditionalFlow(8, 8, 1, 16)
    template = m.TrainState.create(apply_fn=small.apply, params=small.init(0, device="cpu")["params"], tx=m.adamw(1e-4))
    with pytest.raises(ValueError, match="leaves"):
        ck.from_bytes(template, ck.to_bytes(state))
    with pytest.raises(ValueError, match="no 'opt_state'"):
        ck.from_bytes(state, ck.to_bytes({"params": state.params, "step": 0}))

"""world_size-2 gloo test of the data-parallel host logic (the N > 1 path), run on CPU.

Each rank computes the gradient of its batch shard with the oracle (the CUDA step is not available on
CPU; the host logic under test is sharding, the flat-gradient sum all-reduce and the 1/world scaling
applied by the optimiser).  The result must equal the single-process gradient of the whole batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import imf_np

D, L, C, NB, B = 16, 8, 8, 2, 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    p = imf_np.init_params(D, L, C, NB, seed=0, dtype=np.float64, bias_scale=0.1)
    rng = np.random.default_rng(1)
    x, e = rng.standard_normal((B, D)), rng.standard_normal((B, D))
    t = rng.uniform(0.2, 0.9, (B, 1))
    r = t * rng.uniform(0.0, 1.0, (B, 1))
    return p, x, e, t, r


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from meanflow_audio_codec_b200.data_parallel import DataParallel
    dp = DataParallel(backend="gloo")
    p, x, e, t, r = _data()
    a, b = dp.shard_rows(B)
    assert (a, b) == (rank * B // world, (rank + 1) * B // world)
    _, g, _ = imf_np.imf_loss_and_grads(p, x[a:b], e[a:b], t[a:b], r[a:b])
    flat = torch.from_numpy(imf_np.flatten(g, D, L, C, NB).copy())
    dp.allreduce_sum_(flat)
    flat *= 1.0 / dp.world              # what mfac_adamw_step's grad_scale does on the device
    worst = dp.max_over_ranks(float(rank))
    dp.barrier()
    if rank == 0:
        out.put((flat.numpy(), worst))
    dp.destroy()


def test_two_rank_gradient_allreduce_equals_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    flat, worst = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, x, e, t, r = _data()
    _, g, _ = imf_np.imf_loss_and_grads(p, x, e, t, r)
    ref = imf_np.flatten(g, D, L, C, NB)
    np.testing.assert_allclose(flat, ref, rtol=1e-10, atol=1e-14)
    assert worst == 1.0


class _Grads:
    def __init__(self, flat):
        self.flat = flat


class _State:
    applied = None

    def apply_gradients(self, *, grads, grad_scale=1.0):
        self.applied = grads.flat * grad_scale
        return self


class _Strategy:
    """Stands in for ImprovedMeanFlowLoss on CPU: oracle gradient of the shard, handed out bucket by bucket through the
    grad_ready callback exactly like mfac_imf_loss_grad does (blocks last to first, then the encoder)."""

    def compute_loss(self, state, key, x, *, grad_ready=None, row_offset=0):
        p, X, e, t, r = _data()
        a, b = row_offset, row_offset + x.shape[0]
        loss, g, _ = imf_np.imf_loss_and_grads(p, X[a:b], e[a:b], t[a:b], r[a:b])
        flat = torch.from_numpy(imf_np.flatten(g, D, L, C, NB).copy())
        n_blk = (flat.numel() - sum(int(np.prod(s)) for n, s in imf_np.param_shapes(D, L, C, NB) if n.startswith("encoder"))) // NB
        if grad_ready is not None:
            for k in reversed(range(NB)):
                grad_ready(flat[k * n_blk:(k + 1) * n_blk])
            grad_ready(flat[NB * n_blk:])
        return torch.tensor(loss), _Grads(flat)


def _worker_buckets(rank, world, port, out, overlap):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), MFAC_DP_OVERLAP=overlap)
    from meanflow_audio_codec_b200.data_parallel import DataParallel, train_step_dp
    dp = DataParallel(backend="gloo")
    a, b = dp.shard_rows(B)
    state, loss, _ = train_step_dp(dp, _State(), 0, torch.zeros(b - a, D), _Strategy())
    dp.barrier()
    if rank == 0:
        out.put(state.applied.numpy())
    dp.destroy()


@pytest.mark.parametrize("overlap", ["0", "1"])
def test_allreduce_in_train_step_dp(overlap):
    """train_step_dp's N > 1 path, both exchange schedules: one bucket after the backward (default) and per-bucket async
    all-reduce started from the grad_ready callback; waited before the optimiser, 1/world folded into the update."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_buckets, args=(r, world, port, q, overlap)) for r in range(world)]
    for pr in procs:
        pr.start()
    applied = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, x, e, t, r = _data()
    _, g, _ = imf_np.imf_loss_and_grads(p, x, e, t, r)
    np.testing.assert_allclose(applied, imf_np.flatten(g, D, L, C, NB), rtol=1e-10, atol=1e-14)


def test_uneven_global_batch_is_rejected(monkeypatch):
    from meanflow_audio_codec_b200.data_parallel import DataParallel
    monkeypatch.setenv("WORLD_SIZE", "1")
    dp = DataParallel()
    assert dp.shard_rows(10) == (0, 10) and not dp.enabled
    dp.world = 3
    with pytest.raises(ValueError):
        dp.shard_rows(10)

"""Batch-sharded data parallelism for the iMF training step (one process per GPU).

The reference has no distributed code.  The only exchange step of the path is a
sum all-reduce of the flat fp32 gradient; the 1/world factor is folded into the AdamW kernel
(``grad_scale``).  MDCT, encode and sampling shard by rows and need no collective.

``torch.distributed`` is the plumbing (NCCL over NVLink on GPUs, gloo on CPU for tests); the flat
gradient buffer is handed to it as ONE bucket -- NVSwitch all-reduce cost is latency- not link-bound.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, backend: str | None = None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.enabled = self.world > 1
        if self.enabled and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group(backend, rank=self.rank, world_size=self.world,
                                        device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend, rank=self.rank, world_size=self.world)

        self.fused = False

    def init_library_comm(self) -> bool:
        """Second NCCL communicator owned by libmfac (``mfac_comm_init``): the fused training step issues the gradient
        all-reduce itself, slice by slice, on the stream each slice becomes final on.  The 128-byte NCCL id is made by
        rank 0 and broadcast over ``torch.distributed``.  Returns False (and leaves the torch path in charge) on CPU / gloo."""
        if not self.enabled or not torch.cuda.is_available() or dist.get_backend() != "nccl":
            return False
        import ctypes as C

        from . import _lib
        buf = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(_lib.lib().mfac_comm_unique_id(buf), "comm_unique_id")
        idt = torch.tensor(list(buf), dtype=torch.uint8, device=torch.device("cuda", self.local_rank))
        dist.broadcast(idt, src=0)
        raw = (C.c_ubyte * 128)(*idt.cpu().tolist())
        with torch.cuda.device(self.local_rank):
            _lib.check(_lib.lib().mfac_comm_init(raw, self.rank, self.world), "comm_init")
        self.fused = True
        return True

    # ------------------------------------------------------------ sharding
    def shard_rows(self, n: int) -> tuple[int, int]:
        """[start, stop) of this rank's rows of a global batch of n (n must divide evenly: the
        weighted-L2 loss is a batch mean, so equal shards make mean-of-means exact)."""
        if n % self.world:
            raise ValueError(f"global batch {n} is not divisible by world size {self.world}")
        per = n // self.world
        return self.rank * per, (self.rank + 1) * per

    def shard(self, x: torch.Tensor) -> torch.Tensor:
        a, b = self.shard_rows(x.shape[0])
        return x[a:b]

    # ------------------------------------------------------------ collectives
    def allreduce_sum_(self, flat: torch.Tensor) -> torch.Tensor:
        if self.enabled:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return flat

    def barrier(self):
        if self.enabled:
            dist.barrier()

    def max_over_ranks(self, value: float, device=None) -> float:
        if not self.enabled:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def destroy(self):
        if self.fused:
            from . import _lib
            _lib.lib().mfac_comm_destroy()
            self.fused = False
        if self.enabled and dist.is_initialized():
            dist.destroy_process_group()


def train_step_dp(dp: DataParallel, state, key, x_local, loss_strategy, **kw):
    """Data-parallel ``train_step``: local loss/grad -> sum all-reduce of the flat gradient -> AdamW with
    grad_scale = 1/world.  The RNG rows are offset by rank so shards draw independent (e, t, r); the
    "first half gets r = t" rule (utils.py:41-44) is applied per local shard."""
    kw.setdefault("row_offset", dp.rank * x_local.shape[0])
    fused_ok = hasattr(loss_strategy, "train_step_fused") and hasattr(state, "tx")
    if fused_ok and dp.enabled and getattr(dp, "fused", False):
        # Small batches (the library's concurrent schedule): gradient buckets are exchanged over libmfac's own communicator on a
        # side stream under the rest of the backward (2 x B200, 128 rows / GPU: 1.40 ms / step against 1.80 ms).  Large
        # batches keep ONE all-reduce after the backward, issued from here: measured on 8 x B200 at 37 888 rows / GPU the same
        # collective issued from inside the library call cost 16.6 ms / step against 15.1 ms (both communicators move 113 MB
        # in 345 us when timed alone, tools/comm_probe.py; the host thread cannot run ahead of a collective enqueued mid-call).
        import ctypes as C

        from . import _lib
        fused_ok = bool(_lib.lib().mfac_uses_concurrent_schedule(C.byref(state.model.dims), int(x_local.shape[0])))
    if fused_ok and (not dp.enabled or getattr(dp, "fused", False)):
        # ONE library call: loss/grad schedule, bucketed NCCL all-reduce over libmfac's own communicator (world > 1) and AdamW
        state, loss, _ = loss_strategy.train_step_fused(state, key, x_local, world=dp.world if dp.enabled else 1, **kw)
        return state, loss, key
    if not dp.enabled:
        loss, grads = loss_strategy.compute_loss(state, key, x_local, **kw)
    elif os.environ.get("MFAC_DP_OVERLAP", "0") != "1":
        # Default: ONE bucket after the whole backward.  Measured on 8 x B200 (profiles/r01_bench_dp8_*): 10.43 ms/step
        # against 10.67 ms for the overlapped variant below at 18 944 rows/GPU (4.74 vs 4.79 ms at 4096): the persistent
        # one-CTA-per-SM GEMMs lose more from the SMs NCCL's kernels occupy than the exchange (113 MB, ~0.3 ms) costs.
        loss, grads = loss_strategy.compute_loss(state, key, x_local, **kw)
        dp.allreduce_sum_(grads.flat)
    else:
        # MFAC_DP_OVERLAP=1: bucketed, overlapped exchange -- every block's gradient slice is all-reduced (async, on NCCL's
        # stream) as soon as the launches that finalise it are enqueued (grad_ready callback of mfac_imf_loss_grad)
        works = []
        loss, grads = loss_strategy.compute_loss(
            state, key, x_local, grad_ready=lambda sl: works.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, async_op=True)), **kw)
        for w in works:
            w.wait()          # the compute stream waits for the exchange; no host synchronisation
    state = state.apply_gradients(grads=grads, grad_scale=1.0 / dp.world)
    return state, loss, key

#!/usr/bin/env python
"""N-GPU check + timing of the data-parallel step (launch with torchrun): the fused path (mfac_imf_train_step all-reducing
through libmfac's own NCCL communicator, slice by slice) against the torch.distributed path (one all-reduce after the backward).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_check.py [batches...]"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import meanflow_audio_codec_b200 as m
from meanflow_audio_codec_b200.data_parallel import DataParallel, train_step_dp

dp = DataParallel()
torch.cuda.set_device(dp.local_rank)
dev = torch.device("cuda", dp.local_rank)
batches = [int(a) for a in sys.argv[1:]] or [128, 4096]


def fresh(D=1024):
    model = m.ConditionalFlow(noise_dimension=D, condition_dimension=128, num_blocks=8, latent_dimension=256)
    return m.TrainState.create(apply_fn=model.apply, params=model.init(42, device=dev)["params"], tx=m.adamw(1e-4, 1e-4))


strat = m.ImprovedMeanFlowLoss()
ok = True
# ---- parity: same initial state, per-rank shards, torch path vs fused path
x = torch.randn(256, 1024, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + dp.rank))
sa = fresh()
for _ in range(2):
    sa, la, _ = train_step_dp(dp, sa, 0, x, strat)
assert dp.init_library_comm()
sb = fresh()
for _ in range(2):
    sb, lb, _ = train_step_dp(dp, sb, 0, x, strat)
fa, fb = sa.model.flat_params(sa.params).flat, sb.model.flat_params(sb.params).flat
p0 = fresh().model.flat_params(fresh().params).flat
err = float(((fb - fa).norm() / (fa - p0).norm()))
print(f"rank {dp.rank}: fused vs torch DP update rel err {err:.2e}, loss {float(la):.6f} / {float(lb):.6f}", flush=True)
ok &= err < 1e-3 and abs(float(la) - float(lb)) < 1e-5
# all ranks hold identical parameters afterwards
chk = fb[::997].clone()
ref = chk.clone()
torch.distributed.broadcast(ref, src=0)
ok &= bool(torch.equal(chk, ref))

# ---- timing
for fused in (False, True):
    dp.fused = fused
    for B in batches:
        st = fresh()
        xx = torch.randn(B, 1024, device=dev)
        for _ in range(5):
            st, loss, _ = train_step_dp(dp, st, 0, xx, strat)
        torch.cuda.synchronize(); dp.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30 if B <= 4096 else 10
        e0.record()
        for _ in range(n):
            st, loss, _ = train_step_dp(dp, st, 0, xx, strat)
        e1.record()
        torch.cuda.synchronize()
        ms = dp.max_over_ranks(e0.elapsed_time(e1) / n, dev)
        if dp.rank == 0:
            print(f"world={dp.world} fused={int(fused)} B/gpu={B:6d}: {ms:7.3f} ms/step  {B * dp.world / ms * 1e3:10.0f} samples/s", flush=True)
        del st, xx
        torch.cuda.empty_cache()
dp.fused = True
if dp.rank == 0:
    print("DP CHECK", "OK" if ok else "FAILED", flush=True)
dp.destroy()
sys.exit(0 if ok else 1)

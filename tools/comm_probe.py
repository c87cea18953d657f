#!/usr/bin/env python
"""All-reduce of a gradient-sized fp32 buffer: libmfac's own NCCL communicator (mfac_comm_*) against torch.distributed's.
usage: torchrun --nproc-per-node N tools/comm_probe.py"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from meanflow_audio_codec_b200 import _lib
from meanflow_audio_codec_b200.data_parallel import DataParallel

dp = DataParallel()
torch.cuda.set_device(dp.local_rank)
dev = torch.device("cuda", dp.local_rank)
assert dp.init_library_comm()
l = _lib.lib()
for n in (28_262_272, 6_860_544, 1_000_000):
    buf = torch.ones(n, device=dev)
    def t_lib():
        _lib.check(l.mfac_comm_allreduce_sum_f32(buf.data_ptr(), n, _lib.stream_ptr()), "allreduce")
    def t_torch():
        dist.all_reduce(buf)
    for name, fn in (("libmfac", t_lib), ("torch", t_torch), ("libmfac", t_lib), ("torch", t_torch)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = dp.max_over_ranks(e0.elapsed_time(e1) / 20, dev)
        if dp.rank == 0:
            print(f"world={dp.world} n={n:9d} ({n * 4 / 1e6:6.1f} MB) {name:8s}: {ms * 1e3:8.1f} us  busbw {2 * (dp.world - 1) / dp.world * n * 4 / ms / 1e6:7.1f} GB/s", flush=True)
        buf.fill_(1.0)
dp.destroy()

#!/usr/bin/env python
"""MDCT / IMDCT on 10 s clips: timing (CUDA events) or a short run for ncu.  usage: mdct_bench.py [clips] [iters]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import meanflow_audio_codec_b200 as m

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, N, hop = 441000, 512, 256
x = 0.1 * torch.randn(B, T, device="cuda")
X = m.mdct(x, N, hop); y = m.imdct(X, N, hop)
torch.cuda.synchronize()
for name, fn, nbytes in (("mdct", lambda: m.mdct(x, N, hop), 4 * (x.numel() + X.numel())),
                         ("imdct", lambda: m.imdct(X, N, hop), 4 * (X.numel() + y.numel()))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: B={B} {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  {B * 10 / ms * 1e3:.0f} audio-s/s")

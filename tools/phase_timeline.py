#!/usr/bin/env python
"""Phase timeline of one eager training step (debug marks inside libmfac). usage: python tools/phase_timeline.py [B] [conc_rows]"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import meanflow_audio_codec_b200 as m
from meanflow_audio_codec_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
conc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
_lib.set_concurrency_max_rows(conc)
model = m.ConditionalFlow(noise_dimension=1024, condition_dimension=128, num_blocks=8, latent_dimension=256)
state = m.TrainState.create(apply_fn=model.apply, params=model.init(42)["params"], tx=m.adamw(1e-4, 1e-4))
strat = m.ImprovedMeanFlowLoss()
x = torch.randn(B, 1024, device="cuda")
for _ in range(5):
    state, loss, _ = m.train_step(state, 0, x, strat)
torch.cuda.synchronize()
l = _lib.lib()
acc = {}
for rep in range(10):
    l.mfac_debug_phase_marks(1)
    state, loss, _ = m.train_step(state, 0, x, strat)
    l.mfac_debug_phase_marks(0)
    ids, ms = (C.c_int32 * 64)(), (C.c_float * 64)()
    n = l.mfac_debug_phase_collect(ids, ms, 64)
    for i in range(n):
        acc.setdefault(ids[i], []).append(ms[i])
print(f"B={B} conc_rows={conc}: phase id -> ms since step start (median of 10)")
prev = 0.0
for k in sorted(acc):
    v = sorted(acc[k])[len(acc[k]) // 2]
    print(f"  {k:3d}: {v * 1e3:8.1f} us  (+{(v - prev) * 1e3:7.1f})")
    prev = v

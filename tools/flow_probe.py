#!/usr/bin/env python
"""Forward time of the full-size mixer / ConvNeXt velocity network (D = 1024, 8 blocks).
usage (GPU box): python tools/flow_probe.py {mlp_mixer|convnet} [batch]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import meanflow_audio_codec_b200 as m

arch = sys.argv[1] if len(sys.argv) > 1 else "convnet"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (256 if arch == "mlp_mixer" else 2048)
cls = m.ConditionalMLPMixerFlow if arch == "mlp_mixer" else m.ConditionalConvFlow
model = cls(1024, 128, 8, 256)
params = model.init(42)["params"]
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B, 1024, device="cuda", generator=g)
t = torch.rand(B, 2, device="cuda", generator=g)
lat = torch.randn(B, 32, 256, device="cuda", generator=g)
for _ in range(3):
    y = model.apply({"params": params}, x, t, lat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = model.apply({"params": params}, x, t, lat)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{arch} B={B}: {ms:.3f} ms  {B / ms * 1e3:.0f} rows/s  finite={bool(torch.isfinite(y).all())}")

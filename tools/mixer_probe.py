#!/usr/bin/env python
"""Full-size MLP-Mixer forward (D = 1024, 8 blocks, B = 256): timing, and -- run once plainly and once with
MFAC_NO_FUSED_CHANNEL_MIX=1 -- the difference between the fused channel-mix kernel and the two-GEMM path."""
import sys, os
sys.path.insert(0, ".")
import torch, numpy as np
import meanflow_audio_codec_b200 as m
# A/B: fused channel mix vs two GEMMs on the full-size mixer
def run(B=256):
    model = m.ConditionalMLPMixerFlow(1024, 128, 8, 256)
    params = model.init(42)["params"]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, 1024, device="cuda", generator=g); t = torch.rand(B, 2, device="cuda", generator=g)
    lat = torch.randn(B, 32, 256, device="cuda", generator=g)
    y = model.apply({"params": params}, x, t, lat)
    for _ in range(3): y = model.apply({"params": params}, x, t, lat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = model.apply({"params": params}, x, t, lat)
    e1.record(); torch.cuda.synchronize()
    return y, e0.elapsed_time(e1) / 5
y, ms = run()
print("fused" if not os.environ.get("MFAC_NO_FUSED_CHANNEL_MIX") else "two-gemm", "ms", ms, "rows/s", 256 / ms * 1e3, "finite", bool(torch.isfinite(y).all()), "norm", float(y.norm()))
torch.save(y.cpu(), "/tmp/mixer_y_%s.pt" % ("a" if not os.environ.get("MFAC_NO_FUSED_CHANNEL_MIX") else "b"))
if os.path.exists("/tmp/mixer_y_a.pt") and os.path.exists("/tmp/mixer_y_b.pt"):
    a, b = torch.load("/tmp/mixer_y_a.pt"), torch.load("/tmp/mixer_y_b.pt")
    print("rel diff fused vs two-gemm", float((a - b).norm() / b.norm()))

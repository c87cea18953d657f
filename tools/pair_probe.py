#!/usr/bin/env python
"""CTA-pair (cta_group::2) GEMM: correctness against torch for all operand majors and timing against the 1-CTA kernel."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from meanflow_audio_codec_b200 import _lib

lib = _lib.lib()


def run(M, N, K, a_mn, b_mn, pair, iters=0):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(K, N, device="cuda", generator=g).to(torch.bfloat16)
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.contiguous() if b_mn else B.t().contiguous()
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.mfac_debug_set_pair_gemm(1 if pair else 0)
    call = lambda: lib.mfac_debug_gemm_bf16(As.data_ptr(), Bs.data_ptr(), out.data_ptr(), M, N, K, a_mn, b_mn, 0, None)
    rc = call()
    torch.cuda.synchronize()
    ref = A.float() @ B.float()
    err = ((out - ref).norm() / ref.norm()).item()
    ms = None
    if iters:
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    lib.mfac_debug_set_pair_gemm(1)
    return rc, err, ms


ok = True
for (M, N, K) in [(18944, 1280, 1280), (18944, 1024, 1280), (19000, 1280, 1280), (18944, 1280, 520)]:
    for a_mn, b_mn in [(0, 1), (0, 0), (1, 1), (1, 0)]:
        if a_mn and M % 8:
            continue
        rc, err, _ = run(M, N, K, a_mn, b_mn, True)
        print(f"pair M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: rc={rc} rel_err={err:.3e}", flush=True)
        ok &= rc == 0 and err < 1e-5
for (M, N, K, a_mn, b_mn) in [(18944, 1280, 1280, 0, 1), (18944, 1024, 1280, 0, 1), (18944, 1280, 1280, 0, 0), (55104, 1280, 1280, 0, 1)]:
    for pair in (0, 1):
        rc, err, ms = run(M, N, K, a_mn, b_mn, pair, iters=10)
        print(f"time pair={pair} M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s err={err:.2e}", flush=True)
print("PAIR PROBE", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

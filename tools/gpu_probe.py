"""First-contact GPU probe: tcgen05 GEMM (all operand-major combinations, both tile widths)
and the MDCT/IMDCT kernels against torch / the NumPy oracle.  Run under gpurun with a timeout."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
lib = C.CDLL(str(ROOT / "meanflow_audio_codec_b200" / "libmfac.so"))
lib.mfac_status_string.restype = C.c_char_p
P, I64, I32 = C.c_void_p, C.c_int64, C.c_int32
lib.mfac_debug_gemm_bf16.argtypes = [P, P, P, I64, I64, I64, I32, I32, I32, P]
lib.mfac_mdct_f32.argtypes = [P, P, I64, I64, I32, I32, P]
lib.mfac_imdct_f32.argtypes = [P, P, I64, I64, I32, I32, P]
lib.mfac_debug_set_simt_gemm.argtypes = [I32]


def st(code):
    return lib.mfac_status_string(code).decode()


def gemm_case(M, N, K, a_mn, b_mn, bn, simt=False):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(K, N, device="cuda", generator=g).to(torch.bfloat16)
    ref = A.float() @ B.float()
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.contiguous() if b_mn else B.t().contiguous()
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.mfac_debug_set_simt_gemm(1 if simt else 0)
    rc = lib.mfac_debug_gemm_bf16(As.data_ptr(), Bs.data_ptr(), out.data_ptr(), M, N, K, a_mn, b_mn, bn, None)
    torch.cuda.synchronize()
    lib.mfac_debug_set_simt_gemm(0)
    err = ((out - ref).norm() / ref.norm()).item() if rc == 0 else float("nan")
    print(f"gemm M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={bn} simt={simt}: rc={rc} ({st(rc)}) rel_err={err:.3e}",
          flush=True)
    return rc == 0 and err < 1e-5


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    ok &= gemm_case(128, 128, 64, 0, 0, 128, simt=True)
    for a_mn, b_mn in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        for (M, N, K, bn) in [(128, 128, 64, 128), (128, 256, 128, 256), (256, 384, 320, 128), (1000, 1280, 1280, 256),
                              (4096, 1024, 1280, 0), (100, 64, 72, 128)]:
            if a_mn and M % 8:   # an MN-major A has leading dimension M: TMA needs 16-byte row strides
                continue
            try:
                ok &= gemm_case(M, N, K, a_mn, b_mn, bn)
            except Exception as e:  # a trap poisons the context: stop
                print("EXC", e, flush=True)
                return 1
    # timing of the big ones
    for (M, N, K, a_mn, b_mn, bn) in [(18944, 1280, 1280, 0, 1, 256), (18944, 1280, 1280, 0, 1, 128),
                                       (18944, 1024, 1280, 0, 1, 256), (4096, 1280, 1280, 0, 1, 128),
                                       (1280, 1280, 18944, 1, 1, 128)]:
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda")
        for _ in range(3):
            lib.mfac_debug_gemm_bf16(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, a_mn, b_mn, bn, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.mfac_debug_gemm_bf16(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, a_mn, b_mn, bn, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"time M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

    # MDCT
    from oracle import mdct_np
    for (B, T, N, hop) in [(3, 4000, 512, 256), (2, 3000, 512, 512), (1, 1024, 256, 128), (2, 3000, 576, 288), (4, 784, 512, 256),
                           (2, 100, 512, 256), (2, 2000, 512, 100)]:
        x = torch.randn(B, T, generator=torch.Generator().manual_seed(T))
        ref = mdct_np.mdct(x.numpy(), N, hop)
        nf = ref.shape[1]
        xd = x.cuda()
        X = torch.full((B, nf, N), float("nan"), device="cuda")
        rc = lib.mfac_mdct_f32(xd.data_ptr(), X.data_ptr(), B, T, N, hop, None)
        torch.cuda.synchronize()
        err = np.linalg.norm(X.cpu().numpy() - ref) / np.linalg.norm(ref)
        yref = mdct_np.imdct(ref, N, hop)
        Xd = torch.from_numpy(ref.astype(np.float32)).cuda()
        y = torch.full((B, yref.shape[1]), float("nan"), device="cuda")
        rc2 = lib.mfac_imdct_f32(Xd.data_ptr(), y.data_ptr(), B, nf, N, hop, None)
        torch.cuda.synchronize()
        err2 = np.linalg.norm(y.cpu().numpy() - yref) / np.linalg.norm(yref)
        print(f"mdct B={B} T={T} N={N} hop={hop}: rc={rc},{rc2} rel_err fwd={err:.3e} inv={err2:.3e}", flush=True)
        ok &= rc == 0 and rc2 == 0 and err < 1e-5 and err2 < 1e-5
    # MDCT timing: 64 clips of 10 s
    B, T, N, hop = 256, 441000, 512, 256
    x = 0.1 * torch.randn(B, T, device="cuda")
    nf = (T - N) // hop + 1
    X = torch.empty(B, nf, N, device="cuda")
    L = (nf - 1) * hop + 2 * N
    y = torch.empty(B, L, device="cuda")
    for name, fn in [("mdct", lambda: lib.mfac_mdct_f32(x.data_ptr(), X.data_ptr(), B, T, N, hop, None)),
                     ("imdct", lambda: lib.mfac_imdct_f32(X.data_ptr(), y.data_ptr(), B, nf, N, hop, None))]:
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        bytes_ = 4 * B * (T + nf * N) if name == "mdct" else 4 * B * (nf * N + L)
        print(f"time {name} B={B}: {ms:.3f} ms  {bytes_/ms/1e6:.0f} GB/s  {B*10/ms*1e3:.0f} audio-s/s", flush=True)
    # round trip gain
    rt = (y[:, 2 * N:T - 2 * N] - 2 * x[:, 2 * N:T - 2 * N]).norm() / (2 * x[:, 2 * N:T - 2 * N]).norm()
    print("round-trip rel err vs 2x:", rt.item())
    print("PROBE", "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 2


if __name__ == "__main__":
    sys.exit(main())

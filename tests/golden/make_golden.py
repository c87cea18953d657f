"""Generates the committed golden vectors from the REFERENCE's own code, in the build container.

    python tests/golden/make_golden.py        (needs /root/reference; never run on the GPU box)

Source of truth: /root/reference/test/test_mdct_utils.py (``mdct_baseline`` / ``imdct_baseline``,
pure NumPy, fp32) -- the baseline the reference's only MDCT test pins its JAX implementation to
(/root/reference/test/test_mdct.py:13-56).  JAX itself is not installable here, so these are the
only reference-executed numbers available for the path.  Cases:
  g1   test_mdct.py verbatim: np.random.seed(42), x = randn(1024) fp32, N=256, hop=128
  g2.. N=512/hop=256 (every shipped config), hop=N, default N=576, T < N, odd hop, batch 3
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference/test")
from test_mdct_utils import imdct_baseline, mdct_baseline  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    cases = {}
    np.random.seed(42)
    x = np.random.randn(1024).astype(np.float32)
    cases["g1"] = (x[None, :], 256, 128)
    rng = np.random.default_rng(1234)
    cases["g2_n512_h256"] = (rng.standard_normal((3, 3000)).astype(np.float32), 512, 256)
    cases["g3_n512_h512"] = (rng.standard_normal((2, 2600)).astype(np.float32), 512, 512)
    cases["g4_n576_default"] = (rng.standard_normal((2, 2500)).astype(np.float32), 576, None)
    cases["g5_short"] = (rng.standard_normal((2, 100)).astype(np.float32), 512, 256)
    cases["g6_oddhop"] = (rng.standard_normal((2, 1500)).astype(np.float32), 512, 100)
    cases["g7_mnist"] = (rng.uniform(-1, 1, (4, 784)).astype(np.float32), 512, 256)
    cases["g8_n64"] = (rng.standard_normal((2, 700)).astype(np.float32), 64, 16)
    blob = {}
    for name, (x, N, hop) in cases.items():
        X = mdct_baseline(x, N, hop)
        y = imdct_baseline(X, N, hop)
        blob[name + "/x"] = x
        blob[name + "/X"] = X.astype(np.float32)
        blob[name + "/y"] = y.astype(np.float32)
        blob[name + "/cfg"] = np.array([N, -1 if hop is None else hop], dtype=np.int64)
    np.savez_compressed(OUT / "mdct_reference_baseline.npz", **blob)
    print("wrote", OUT / "mdct_reference_baseline.npz", sum(v.nbytes for v in blob.values()) / 1e6, "MB raw")


def copy_reference_config():
    """tests/golden/config_imf_mlp_mnist_mdct.json is BASELINE.json configs[0] as the reference ships it
    (configs/method=improved_mean_flow--architecture=mlp--dataset=mnist--tokenization=mdct.json): a data fixture in the
    reference's own file format, used by the drop-in tests (tests/test_abi_cpu.py, tests/test_imf_gpu.py)."""
    import json
    src = Path("/root/reference/configs/method=improved_mean_flow--architecture=mlp--dataset=mnist--tokenization=mdct.json")
    dst = Path(__file__).resolve().parent / "config_imf_mlp_mnist_mdct.json"
    dst.write_text(json.dumps(json.loads(src.read_text()), indent=2, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
    copy_reference_config()


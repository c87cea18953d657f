"""Builds ``libmfac_jax_ffi.so`` (needs jaxlib's XLA FFI headers; not available in the build image)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent


def build() -> Path:
    try:
        import jax.ffi
    except Exception as e:  # noqa: BLE001
        raise SystemExit(f"jax.ffi is not importable here ({e}); build this adapter on a box with jax >= 0.4.38") from e
    out = HERE / "libmfac_jax_ffi.so"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{jax.ffi.include_dir()}", f"-I{PKG.parent / 'include'}",
           "-I/usr/local/cuda/include", str(HERE / "mfac_jax_ffi.cc"), f"-L{PKG}", "-l:libmfac.so",
           "-Wl,-rpath,$ORIGIN/..", "-o", str(out)]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build())
    sys.exit(0)

#!/bin/bash
# build libmfac.so from the repo root regardless of the caller's cwd; non-zero exit on failure
set -euo pipefail
cd "$(dirname "$0")/.."
python -m meanflow_audio_codec_b200.build "$@"

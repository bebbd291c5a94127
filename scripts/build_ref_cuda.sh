#!/bin/bash
# The GPU bar (bench.py --impl reference_cuda): the UNMODIFIED reference package with its CUDA extension
# compiled for sm_100, installed into baseline/_ref (git-ignored, travels to the GPU box).
# /root/reference is read-only and setup.py writes into the source tree, so the install runs from a
# /tmp copy.  About 3-4 minutes on 8 cores (256+ KNN template instantiations).
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/ref_cuda.XXXXXX)"
cp -r "$SRC" "$TMP/reference"
chmod -R u+w "$TMP/reference"
rm -rf "$REPO/baseline/_ref"
mkdir -p "$REPO/baseline/_ref"
FORCE_CUDA=1 TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS="${MAX_JOBS:-8}" \
  python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$REPO/baseline/_ref" "$TMP/reference"
rm -rf "$TMP"
ls "$REPO/baseline/_ref/pytorch3d_pointops"

#!/bin/bash
# Build libwgrt.so from the repo root; print register/spill info of the walk kernels; fail loudly.
set -euo pipefail
cd "$(dirname "$0")/.."
python -m gpu_ray_tracing_for_waveguide_based_ar_display_b200.csrc.build --verbose "$@" 2>&1 | tee /tmp/wgrt_build.log | grep -iE "error|walk_fast_kernelILb0ELb" -A2 | grep -iE "error|Used|spill" || true
if grep -qiE "error" /tmp/wgrt_build.log; then echo "BUILD FAILED"; exit 1; fi
ls -la gpu_ray_tracing_for_waveguide_based_ar_display_b200/libwgrt.so

#!/bin/bash
# compute-sanitizer over small parity cases of the scan kernels and the round-2 kernels (racecheck = shared-memory hazards
# inside a kernel; memcheck = out-of-bounds / misaligned accesses).  The backward runs in per-group mode: the sanitizer
# serialises kernels, the persistent launch needs its recompute kernels beside it.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTT_B200_PERSISTENT=0
( timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_mlp_backward.py tests/test_gpu_mlp_forward.py -q -x \
    -k "test_backward_matches_autograd_of_eager and (1-2-3-2 or 2-2-7-3) or test_forward_matches_oracle" 2>&1 | tail -25 ) > gpurun_out/r02_sanitizer_racecheck.log
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_adaln.py tests/test_gpu_mlp_backward.py tests/test_gpu_full_shape.py -q -x \
    -k "test_ln_affine or 2-2-7-3 or 43_argument or 37-16" 2>&1 | tail -25 ) > gpurun_out/r02_sanitizer_memcheck.log
tail -6 gpurun_out/r02_sanitizer_racecheck.log; tail -6 gpurun_out/r02_sanitizer_memcheck.log

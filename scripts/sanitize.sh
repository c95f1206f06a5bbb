#!/bin/bash
# compute-sanitizer over the library's kernels (race / memory checking).  memcheck on the IPR-op
# parity tests and the smoke step (every kernel family is launched at least once); racecheck on the shared-memory
# heavy SSIM / PDQ / trigger kernels.  Slow (10-50x): sizes in these tests are small.  Output: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_ipr_ops.py -x -q -k "not big" > gpurun_out/sanitize_memcheck_ops.log 2>&1
echo "memcheck ops rc=$?"; tail -3 gpurun_out/sanitize_memcheck_ops.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -3 gpurun_out/sanitize_memcheck_smoke.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_ipr_ops.py -x -q -k "ssim or paste or hash or sign" > gpurun_out/sanitize_racecheck_ops.log 2>&1
echo "racecheck ops rc=$?"; tail -3 gpurun_out/sanitize_racecheck_ops.log

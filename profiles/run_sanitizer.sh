#!/bin/bash
# Race / memory checks of our kernels on a B200 (the reference has none).  Run under gpurun:
#   gpurun --timeout 900 -- bash profiles/run_sanitizer.sh
# Small shapes: compute-sanitizer slows kernels down by 10-100x.
set -x
export ODB_LOGLEVEL=WARNING
T="python -m pytest tests/test_kernels_gpu.py -x -q -k 'embedding or rmsnorm or rope or swiglu or adamw or nesterov or 1024-'"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 bash -c "$T" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 bash -c "$T" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log

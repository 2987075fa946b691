#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "fused_zero or ckpt_resume_data_parallel" > gpurun_out/o_tests.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/o_tests.log | cut -c1-300; grep -E "^E .*(loss at step|Loss at step|FAIL)" gpurun_out/o_tests.log | head -5 | cut -c1-200

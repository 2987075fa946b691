#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for g in 16 8 32 4; do ODB_GEMM_GROUP_M=$g timeout 200 python profiles/gemm_raster_bench.py 2>&1 | tail -9; done > gpurun_out/m_raster.log
for g in 0 4; do ODB_LCE_DX_GROUP_M=$g timeout 200 python profiles/gemm_raster_bench.py 2>&1 | grep "LCE dX"; done >> gpurun_out/m_raster.log
cat gpurun_out/m_raster.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -s 2>&1 | tail -14
sed -i 's/r2c_/r2d_/g' tools/gpu_r2c.sh; bash tools/gpu_r2c.sh

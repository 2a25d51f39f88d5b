#!/bin/bash
# first GPU contact of the tile-streaming kernel: bounded, so that a hang cannot hold the box
set -o pipefail
timeout -s KILL 300 python -m pytest tests/test_gpu_weakloss.py -x -q 2>&1 | tail -30 > gpurun_out/r2_weak.log
cat gpurun_out/r2_weak.log

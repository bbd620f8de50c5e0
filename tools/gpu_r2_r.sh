#!/bin/bash
# round 2, run R: the cooperative small-population kernel K1c -- parity against K1, then the latency sweep
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plugin.py -q -m gpu -x -k "cooperative or block_choice" 2>&1 | tail -5 | tee gpurun_out/r_tests.txt
python tools/small_n_sweep.py 2>&1 | tee gpurun_out/r_small_n_sweep.txt

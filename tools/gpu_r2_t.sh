#!/bin/bash
# round 2, run T: K1c with 4 / 8 warps per CTA -- parity, phase timeline (timing build), latency A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plugin.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/t_tests.txt
for w in 8 4; do echo "NW=$w"; NPLANE_COOP_WARPS=$w NPLANE_LIB=$PWD/build/ab/libnplane_timing.so python tools/k1c_phases.py 3000; done 2>&1 | tee gpurun_out/t_phases.txt
for w in 8 4 0; do NPLANE_COOP_WARPS=$w python tools/k1c_ab.py | sed "s/^/NW=$w /"; done 2>&1 | tee gpurun_out/t_ab.txt
NPLANE_LIB=$PWD/build/ab/libnplane_k1c_v2.so python tools/k1c_ab.py | tee -a gpurun_out/t_ab.txt

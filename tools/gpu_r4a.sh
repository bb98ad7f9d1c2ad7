#!/bin/bash
# round-2 session-4 run: GPU tests, then A/B of the overlapped moment reduction and the batched per-lane search
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/r4a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r4a_pytest.log
CFG="M3DREG_NEQ_EARLY=0;M3DREG_NEQ_EARLY=1;M3DREG_LIB_PATH=build_variants/libm3dreg_nnq_old.so"
bash tools/gpu_ab.sh r4a "$CFG" "--slam none"
bash tools/gpu_ab.sh r4ac1 "$CFG" "--slam none --workload c1"
bash tools/gpu_slamtune.sh "M3DREG_X=0;M3DREG_LIB_PATH=build_variants/libm3dreg_nnq_old.so"

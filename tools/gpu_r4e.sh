#!/bin/bash
# two-level hull enumeration: parity first, then the scatter criterion on C1 / C2 / the C4 sweep
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/r4e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r4e_pytest.log
CFG="M3DREG_X=0;M3DREG_NN_HULL_MIN=512 M3DREG_NN_HULL_RATIO=32;M3DREG_NN_HULL_MIN=2048 M3DREG_NN_HULL_RATIO=64;M3DREG_NN_HULL_MIN=8192 M3DREG_NN_HULL_RATIO=512"
bash tools/gpu_ab.sh r4ec1 "$CFG" "--slam none --workload c1"
bash tools/gpu_ab.sh r4e "$CFG" "--slam none"
bash tools/gpu_slamtune.sh "M3DREG_X=0;M3DREG_NN_SWEEP_HULL_MIN=2048 M3DREG_NN_SWEEP_HULL_RATIO=64;M3DREG_NN_SWEEP_HULL_MIN=8192 M3DREG_NN_SWEEP_HULL_RATIO=512"

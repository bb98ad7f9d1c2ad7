#!/bin/bash
# static slabs of chunks per search warp (L1 locality) and larger search blocks
CFG="M3DREG_NN_STATIC_PCT=0;M3DREG_NN_STATIC_PCT=50;M3DREG_NN_STATIC_PCT=75;M3DREG_NN_STATIC_PCT=90;M3DREG_LIB_PATH=build_variants/libm3dreg_t128.so;M3DREG_LIB_PATH=build_variants/libm3dreg_t128.so M3DREG_NN_STATIC_PCT=75;M3DREG_LIB_PATH=build_variants/libm3dreg_t256.so M3DREG_NN_STATIC_PCT=75"
bash tools/gpu_ab.sh r4b "$CFG" "--slam none"
bash tools/gpu_ab.sh r4bc1 "M3DREG_NN_STATIC_PCT=0;M3DREG_NN_STATIC_PCT=75" "--slam none --workload c1"
bash tools/gpu_slamtune.sh "M3DREG_NN_STATIC_PCT=0;M3DREG_NN_STATIC_PCT=75"

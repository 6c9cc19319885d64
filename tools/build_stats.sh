#!/usr/bin/env bash
# Developer build with -DMMD_STATS (per-tile clock64 timeline of block 0 + whole-launch phase counters of the fused kernel)
# next to the product library (it travels to the GPU box, *.so is git-ignored): mmd_retrieval/dev/libmmd_stats.so.
# Use with MMD_LIB_PATH=<that file> python tools/trace_run.py ...   The product library is not touched.
set -e
cd "$(dirname "$0")/.."
D=${MMD_DEV_LIBS:-multimodal-misinformation-detection_b200/mmd_retrieval/dev}   # developer builds (tools/build_stats.sh; a round-1 libmmd_r1.so for A/B sweeps), shipped to the box only when placed there
mkdir -p $D
MMD_STATS=1 MMD_BUILD_OUT=$D/libmmd_stats.so python multimodal-misinformation-detection_b200/build.py --force
ls -la $D

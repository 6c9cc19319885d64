#!/usr/bin/env bash
# Developer build with -DMMD_STATS (per-tile clock64 timeline of block 0 + whole-launch phase counters of the fused kernel)
# next to the product library (it travels to the GPU box, *.so is git-ignored): mmd_retrieval/dev/libmmd_stats.so.
# Use with MMD_LIB_PATH=<that file> python tools/trace_run.py ...   The product library is not touched.
set -e
cd "$(dirname "$0")/.."
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
mkdir -p $D
MMD_STATS=1 MMD_BUILD_OUT=$D/libmmd_stats.so python multimodal-misinformation-detection_b200/build.py --force
ls -la $D

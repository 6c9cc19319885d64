#!/usr/bin/env bash
# Developer builds with -DMMD_STATS (per-tile clock64 timeline of block 0) next to the product library (they travel to the
# GPU box, *.so is git-ignored):  mmd_retrieval/dev/libmmd_stats_w64.so, ..._w32.so (filter pass with 64 / 32-column TMEM loads).
# Use with MMD_LIB_PATH=<that file> python tools/trace_run.py ...   The product library is not touched.
set -e
cd "$(dirname "$0")/.."
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
mkdir -p $D
MMD_STATS=1 MMD_DEFINES=MMD_FILTER_W=64 MMD_BUILD_OUT=$D/libmmd_stats_w64.so python multimodal-misinformation-detection_b200/build.py --force &
sleep 1
wait
MMD_STATS=1 MMD_DEFINES=MMD_FILTER_W=32 MMD_BUILD_OUT=$D/libmmd_stats_w32.so python multimodal-misinformation-detection_b200/build.py --force
ls -la $D

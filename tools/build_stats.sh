#!/usr/bin/env bash
# Developer build with -DMMD_STATS (per-tile clock64 timeline of block 0) -> tools/_dev/libmmd_stats.so,
# then restore the product library.  Use with MMD_LIB_PATH=tools/_dev/libmmd_stats.so python tools/trace_run.py ...
set -e
cd "$(dirname "$0")/.."
MMD_STATS=1 python multimodal-misinformation-detection_b200/build.py --force
mkdir -p tools/_dev
cp multimodal-misinformation-detection_b200/mmd_retrieval/libmmd.so tools/_dev/libmmd_stats.so
python multimodal-misinformation-detection_b200/build.py --force

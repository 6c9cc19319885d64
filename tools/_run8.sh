# 8-GPU evidence: >= 2-GPU parity tests at world 8, then the bench lines of the configs defined on 8 GPUs.
set -u
export MMD_NO_AUTOBUILD=1
bash tools/gpu_multi_check.sh r4a 8 c5 c4

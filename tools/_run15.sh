export MMD_NO_AUTOBUILD=1
python tools/_dbg_eval.py 2>&1 | tail -25
MMD_LEVELS=0 python tools/_dbg_eval.py 2>&1 | tail -8
for i in 1 2 3; do python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "text_topk_accuracy or text_search_drop_in" 2>&1 | tail -2; done

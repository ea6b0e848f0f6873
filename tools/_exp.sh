timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { echo "=== $*"; env "$@" timeout 120 python tools/dev_time.py 4096 12 2>&1 | grep -E "ms/env-step"; }
run FE_X=base
run FE_HEAVY_K=6
run FE_HEAVY_K=8
run FE_HEAVY_SHIFT=14
run FE_HEAVY_SHIFT=22
run FE_PRED_DECAY=70
run FE_PRED_DECAY=95
run FE_HEAVY_K=8 FE_HEAVY_SHIFT=14
run FE_X=base2

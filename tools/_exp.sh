set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { echo "=== $*"; env "$@" timeout 120 python tools/dev_time.py 4096 8 2>&1 | grep -E "ms/env-step|mean ncon|barrier_wait|total mean"; }
run FE_X=0
run FE_MAXCON=39 FE_WPB=7 FE_HEAVY_K=0
run FE_MAXCON=39 FE_WPB=7 FE_HEAVY_K=4
run FE_MAXCON=39 FE_WPB=7 FE_HEAVY_K=3 FE_HEAVY_SHIFT=24
run FE_MAXCON=39
run FE_MAXCON=36 FE_WPB=5 FE_HEAVY_K=0

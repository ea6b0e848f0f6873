#!/bin/bash
# On the GPU box (under gpurun): launch list of a short bench run + one full capture of the step kernel.
# Results land in gpurun_out/; tools/ncu_extract.py (run in the CPU container) turns the report into profiles/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fe_env_step_kernel -s 8 -c 1 -o gpurun_out/prof_$1 -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$1.log 2>&1
tail -2 gpurun_out/prof_$1.log

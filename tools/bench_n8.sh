P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$P --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2d_n8.log 2>&1; tail -1 gpurun_out/bench_r2d_n8.log > gpurun_out/bench_r2d_n8.json; cut -c1-300 gpurun_out/bench_r2d_n8.json
$P --master-port 29512 bench.py --gpus 8 --agent Baxter --furniture chair_ingolf_0650 --steps 10 --warmup 3 > gpurun_out/bench_r2d_baxter_n8.log 2>&1; tail -1 gpurun_out/bench_r2d_baxter_n8.log > gpurun_out/bench_r2d_baxter_n8.json; cut -c1-300 gpurun_out/bench_r2d_baxter_n8.json
$P --master-port 29513 bench.py --gpus 8 --furniture mixed --envs-per-gpu 8192 --steps 10 --warmup 3 > gpurun_out/bench_r2d_mixed_n8.log 2>&1; tail -1 gpurun_out/bench_r2d_mixed_n8.log > gpurun_out/bench_r2d_mixed_n8.json; cut -c1-300 gpurun_out/bench_r2d_mixed_n8.json
grep -h -i -E "error|Traceback" -A3 gpurun_out/bench_r2d_*.log | head -30

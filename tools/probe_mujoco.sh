mkdir -p gpurun_out
{
echo "== import probes"; for m in mujoco mujoco_py dm_control gym gymnasium pybullet; do python -c "import $m; print('$m', $m.__version__)" 2>&1 | tail -1; done
echo "== pip install probes"; timeout 60 python -m pip install mujoco 2>&1 | tail -3
timeout 60 python -m pip download mujoco-py 2>&1 | tail -2
echo "== find"; find / \( -iname "*mujoco*" -o -iname "libmujoco*" \) -not -path "/proc/*" 2>/dev/null | head
ls /opt/wheelhouse 2>/dev/null | grep -i -E "mujoco|gym|bullet" 
echo "== nproc"; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
} > gpurun_out/probe_mujoco.log 2>&1
cat gpurun_out/probe_mujoco.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_head.log 2>&1; tail -1 gpurun_out/bench_r2_head.log

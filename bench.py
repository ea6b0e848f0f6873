#!/usr/bin/env python
"""Benchmark of the FurnitureEnv.step() hot path (BASELINE.json metric: aggregate env-steps/s, FurnitureSawyerEnv +
table_lack_0825, 4096 envs per GPU; one env-step = one env.step() = 50 mj_steps + action mapping + connect check +
obs + reward).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA engine (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement of the reference loop (oracle/ref_env.py
                                                           over oracle/fe_oracle.c) on all host cores; rank 0 only

Prints ONE JSON line (see the contract in the task statement / DESIGN.md "Measurement").
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aggregate env-steps/sec, Sawyer+table_lack @4096 envs/GPU"
UNIT = "env-steps/s"
ENVS_PER_GPU = 4096
B_ENV = 58700  # algorithmic bytes per env-step, SURVEY.md 8d: 50 * 4*(2 nq + 5 nv + nu) + 4*(obs_dim + act_dim) + 8
WORKLOAD = "FurnitureSawyerEnv + table_lack_0825, control_type=impedance, 50 mj_steps per env-step, random actions U(-1,1)"


def b_env(model, obs_dim, act_dim, nsub=50):
    """B_sub = 4 (2 nq + 5 nv + nu) per mj_step; B_env = nsub B_sub + 4 (obs + act) + 8"""
    return nsub * 4 * (2 * model.nq + 5 * model.nv + model.nu) + 4 * (obs_dim + act_dim) + 8


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return {"hbm_gbs": 6650.0}, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_env_rate(seconds, seed=0):
    """env-steps/s of the CPU oracle env on ONE core for about `seconds` of work (reset excluded, like fps.py:119-127)"""
    import numpy as np

    from furniture_b200 import mjcf
    from oracle.ref_env import OracleFurnitureEnv

    m = mjcf.load_scene("Sawyer", "table_lack_0825")
    env = OracleFurnitureEnv(m)
    env.reset()
    rng = np.random.RandomState(seed)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, done, _ = env.step(rng.uniform(-1, 1, env.dof))
        if done:
            env.reset()
        n += 1
    return n / (time.perf_counter() - t0), n


def _ref_worker(args):
    seed, nsteps = args
    import numpy as np

    from furniture_b200 import mjcf
    from oracle.ref_env import OracleFurnitureEnv

    global _REF_ENV
    if "_REF_ENV" not in globals():
        m = mjcf.load_scene("Sawyer", "table_lack_0825")
        _REF_ENV = OracleFurnitureEnv(m)
        _REF_ENV.cfg.seed = 123 + seed
        _REF_ENV.reset()
        _REF_ENV._rng_act = np.random.RandomState(seed)
    env = _REF_ENV
    for _ in range(nsteps):
        _, _, done, _ = env.step(env._rng_act.uniform(-1, 1, env.dof))
        if done:
            env.reset()
    return nsteps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    chunk = 10  # env.step() calls per worker per bench "step"
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, [(i, chunk) for i in range(cores)], chunksize=1)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            total += sum(pool.map(_ref_worker, [(i, chunk) for i in range(cores)], chunksize=1))
        dt = time.perf_counter() - t0
    value = total / dt
    sample = "%d processes x %d env.step() per bench step (spawned workers keep their env alive; reset excluded)" % (cores, chunk)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": cores, "note": "CPU restatement of the reference loop (mujoco-py/MuJoCo 2.0 absent): oracle/ref_env.py over oracle/fe_oracle.c; "
                   "published anchor 225 env-steps/s on one Xeon 6154 core (docs/more_info.md:35)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_ours(args):
    import numpy as np
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from furniture_b200.env import BatchedFurnitureEnv, ShardedFurnitureEnv

    n_local = args.envs_per_gpu
    if world > 1:
        env = ShardedFurnitureEnv(n_local, furniture_name=args.furniture)
        benv = env.env
    else:
        env = benv = BatchedFurnitureEnv("Sawyer", args.furniture, n_local, device=local, seed=123)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(rank)
    K, W = args.steps, args.warmup
    acts = [torch.rand((n_local, benv.act_dim), device=dev, generator=gen) * 2 - 1 for _ in range(K + W)]
    if args.actions == "settled":  # SURVEY.md 8d "settled" variant: zero arm action, gripper open, no connect request
        for a in acts:
            a.zero_()
            a[:, -2:] = -1.0
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > L2 (126 MB)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        env.step(acts[i])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches = 0
    for k in range(K):
        flush.fill_(float(k))  # evict L2 between timed iterations (not timed)
        ev[k][0].record()
        env.step(acts[W + k])
        ev[k][1].record()
        launches += 2  # fe_env_step_kernel + fe_order_kernel (block packing for the next step)
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = n_local * world * K / (total_ms * 1e-3)

    # end to end through the public API with HOST buffers: pinned actions -> H2D -> step (-> all_gather) -> D2H results
    Ke = max(3, min(K, 10))
    a_host = [torch.rand((n_local, benv.act_dim)).mul_(2).sub_(1).pin_memory() for _ in range(Ke)]
    obs_host = torch.empty((n_local * world if world > 1 else n_local, benv.obs_dim)).pin_memory()
    rew_host = torch.empty(n_local * world if world > 1 else n_local).pin_memory()
    done_host = torch.empty(n_local * world if world > 1 else n_local, dtype=torch.bool if world > 1 else torch.uint8).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        od, rew, done, _ = env.step(a_host[k])
        obs_host[:, : benv.object_ob_dim].copy_(od["object_ob"], non_blocking=True)
        obs_host[:, benv.object_ob_dim :].copy_(od["robot_ob"], non_blocking=True)
        rew_host.copy_(rew, non_blocking=True)
        done_host.copy_(done, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = n_local * world * Ke / float(e2e_s.item())
    h2d = n_local * benv.act_dim * 4
    d2h = obs_host.numel() * 4 + rew_host.numel() * 4 + done_host.numel()

    if rank == 0:
        peaks, peak_src = load_peaks()
        kernel_ms = total_ms / K  # one env-step = one launch of fe_env_step_kernel (+ the all_gather when N > 1)
        benv_bytes = b_env(benv.model, benv.obs_dim, benv.act_dim)
        assert args.furniture != "table_lack_0825" or benv_bytes == B_ENV
        achieved = benv_bytes * n_local / (kernel_ms * 1e-3) / 1e9
        workload = WORKLOAD if (args.furniture == "table_lack_0825" and args.actions == "random") else (
            "FurnitureSawyerEnv + %s, control_type=impedance, 50 mj_steps per env-step, %s" % (args.furniture, "random actions U(-1,1)" if args.actions == "random" else "settled (zero arm action, gripper open)"))
        default_case = args.furniture == "table_lack_0825" and args.actions == "random" and n_local == ENVS_PER_GPU
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and default_case:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        out = {
            "metric": METRIC if default_case else "aggregate env-steps/sec, Sawyer+%s @%d envs/GPU (%s actions)" % (args.furniture, n_local, args.actions),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "envs_per_gpu": n_local, "global_envs": n_local * world, "parallelism": "env-shards x%d" % world,
                       "l2": "flushed between timed steps (256 MiB write)", "timing": "CUDA events per step on the launch stream, max over ranks"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "kernel": "fe_env_step_kernel", "algorithmic_bytes_per_launch": benv_bytes * n_local, "peak_source": peak_src,
                         "note": "state stays in shared memory for the 50 mj_steps of a launch; the path is latency/issue bound, not HBM bound (DESIGN.md)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline and default_case:
            v, n = cpu_env_rate(args.cpu_seconds)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": "%d env.step() of one CPU oracle env (oracle/ref_env.py over oracle/fe_oracle.c) in %.0f s" % (n, args.cpu_seconds)}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--furniture", default="table_lack_0825", help="other furniture = parity-test configs timed for DESIGN.md, not the bench line")
    ap.add_argument("--actions", default="random", choices=["random", "settled"])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Multi-process host logic of the env sharding (world_size 2, gloo, CPU): each rank steps its own shard and one
all_gather makes the packed [obs | reward | done] of every shard visible everywhere.  The shard itself
is the lane-emulated engine here; on the GPU box the same ShardedFurnitureEnv wraps the CUDA engine over NCCL."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from collections import OrderedDict
from furniture_b200 import mjcf
from furniture_b200.env import ShardedFurnitureEnv
from parity_util import make_engine

class EmuShard:  # stand-in for BatchedFurnitureEnv with CPU tensors
    def __init__(self, n, seed):
        self.model = mjcf.load_scene("Sawyer", "table_lack_0825")
        self.engine = make_engine(self.model, n, False, seed=seed, nsub=2)
        self.num_envs, self.obs_dim, self.act_dim = n, self.engine.obs_dim, self.engine.act_dim
        self.object_ob_dim = 7 * self.engine.scene.npart
        self.device = torch.device("cpu")
        self._obs = torch.zeros((n, self.obs_dim))
    def _obs_dict(self, obs):
        return OrderedDict(object_ob=obs[:, : self.object_ob_dim], robot_ob=obs[:, self.object_ob_dim :])
    def reset(self):
        self.engine.env_reset(); self._obs.copy_(torch.from_numpy(self.engine.get("obs"))); return self._obs_dict(self._obs)
    def step(self, a):
        obs, rew, done, info = self.engine.env_step_host(a.numpy())
        self._obs.copy_(torch.from_numpy(obs))
        return self._obs_dict(self._obs), torch.from_numpy(rew), torch.from_numpy(done), torch.from_numpy(info)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 3
shard = EmuShard(n, ShardedFurnitureEnv.shard_seed(123, rank, n))
env = ShardedFurnitureEnv(n, env=shard)
od = env.reset()
assert od["object_ob"].shape == (n * world, 35) and od["robot_ob"].shape == (n * world, 29)
mine = torch.cat([od["object_ob"], od["robot_ob"]], 1)[rank * n : (rank + 1) * n]
assert torch.equal(mine, shard._obs)                        # my slice of the gathered tensor is my shard
other = torch.cat([od["object_ob"], od["robot_ob"]], 1)[(1 - rank) * n : (2 - rank) * n]
assert not torch.allclose(other[:, :2], mine[:, :2])        # different seeds -> different placements
a = torch.rand((n, shard.act_dim), generator=torch.Generator().manual_seed(rank)) * 2 - 1
od, rew, done, info = env.step(a)
assert rew.shape == (n * world,) and done.shape == (n * world,) and done.dtype == torch.bool
full = [torch.zeros((n, shard.obs_dim)) for _ in range(world)]
dist.all_gather(full, shard._obs)
assert torch.equal(torch.cat([od["object_ob"], od["robot_ob"]], 1), torch.cat(full, 0))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_env_shards_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-3000:])
        assert "rank %d ok" % r in o

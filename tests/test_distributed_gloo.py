"""N>1 path on CPU: world_size-2 gloo.  The env batch shards with no data-path collective; the only
collective is the final all-gather of per-env results (BatchedEnv.gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_model


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mujoco_template as mt
    from oracle.oracle import OracleModel

    model = load_model("cartpole")
    lo, hi = mt.shard_range(total, rank, world)
    rng = np.random.default_rng(0)                     # same stream on every rank: global state, local slice
    qpos = rng.uniform(-0.2, 0.2, (total, 2)); qvel = rng.uniform(-0.5, 0.5, (total, 2))
    q, v, u = qpos[lo:hi].copy(), qvel[lo:hi].copy(), np.zeros((hi - lo, 1))
    om = OracleModel(model.blob, dict(nq=2, nv=2, nu=1, nbody=model.nbody, njnt=2, ngeom=model.ngeom, nsite=1, ntendon=0))
    om.batch_rollout(q, v, u, nsteps=20, nthreads=1)   # each rank advances only its shard; nothing is exchanged
    local = torch.as_tensor(q.T.copy())                # (nq, n_local), env axis last
    gathered = mt.BatchedEnv.gather(None, local)       # the path's only collective
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("total", [16, 17])   # 17: shard_range hands out blocks of 8 and 9 envs (ragged gather)
def test_two_rank_sharded_rollout_matches_single_process(tmp_path, total):
    world = 2
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    gathered = torch.load(out).numpy()
    from oracle.oracle import OracleModel

    model = load_model("cartpole")
    rng = np.random.default_rng(0)
    qpos = rng.uniform(-0.2, 0.2, (total, 2)); qvel = rng.uniform(-0.5, 0.5, (total, 2))
    om = OracleModel(model.blob, dict(nq=2, nv=2, nu=1, nbody=model.nbody, njnt=2, ngeom=model.ngeom, nsite=1, ntendon=0))
    om.batch_rollout(qpos, qvel, np.zeros((total, 1)), nsteps=20, nthreads=1)
    assert gathered.shape == (2, total)
    assert np.array_equal(gathered, qpos.T)            # bit-identical: sharding does not change any env's arithmetic

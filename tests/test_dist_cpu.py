"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: parameter broadcast, camera scatter, frame sharding and
statistics gather (autovfx_b200/render_loop.py).  The same code runs over NCCL on the GPUs (bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autovfx_b200 import render_loop as RL
from autovfx_b200 import scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, mode, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        g = scene.synthetic_gaussians(500, seed=9) if rank == 0 else None
        gg = RL.broadcast_gaussians(g, dev)
        ref = scene.synthetic_gaussians(500, seed=9)
        ok_params = all(torch.equal(gg[k], ref[k]) for k in ref)
        cams = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=n_frames, w=64, h=48))
        packed = RL.pack_cameras(cams)
        mine = RL.scatter_cameras(packed if rank == 0 else None, n_frames, dev, mode=mode)
        idx = RL.shard_indices(n_frames, rank, world, mode)
        ok_cams = mine.shape == (len(idx), RL.CAM_FLOATS) and torch.equal(mine, packed[idx])
        # each rank "renders" its frames: the statistic is the frame index, so rank 0 can check full coverage
        local = torch.full((max(1, (n_frames + world - 1) // world),), -1.0)
        local[:len(idx)] = torch.tensor(idx, dtype=torch.float32)
        gathered = RL.gather_stats(local)
        ok_gather = True
        if rank == 0:
            seen = sorted(int(v) for t in gathered for v in t.tolist() if v >= 0)
            ok_gather = seen == list(range(n_frames))
        ret[rank] = bool(ok_params and ok_cams and ok_gather)
    finally:
        dist.destroy_process_group()


def _run(n_frames, mode):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, mode, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_scatter_broadcast_gather_roundrobin():
    _run(9, "roundrobin")


def test_scatter_broadcast_gather_block_uneven():
    _run(5, "block")

"""SURVEY §8d config 5 (stand-in: the SuGaR checkpoint is not available offline): a 5M-Gaussian scene stored SuGaR-style
(M=25 SH coefficients, 300-byte rows) plus one inserted object of 200k Gaussians that moves rigidly every frame, rendered at
1920x1080 as the reference's frame loop does — per frame: object edit + merge, render() (SH image + normal image + normal
maps), 8-bit conversions, hand-off to the host — through ResidentScene + FrameLoop(product=True, pack8=True).

    python tools/bench_config5.py [--frames 150] [--scene 5000000] [--object 200000]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config5.py   # frames sharded over ranks

Prints one JSON line (rank 0).  Reference behaviour reproduced: merged models render with SH degree 0 (gaussians_utils.py:75 builds
a fresh GaussianModel whose active_sh_degree is 0); --sh-degree 3 renders the trained degree instead.
"""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovfx_b200 import edit, scene  # noqa: E402
from autovfx_b200 import render_loop as RL  # noqa: E402


def to_raw(g):
    M = g["shs"].shape[1]
    op = g["opacities"].clamp(1e-6, 1 - 1e-6)
    return {"xyz": g["means3D"], "f_dc": g["shs"][:, :1].contiguous(), "f_rest": g["shs"][:, 1:].contiguous(),
            "opacity": torch.log(op / (1 - op)), "scaling": torch.log(g["scales"]), "rotation": g["rotations"]}, M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--scene", type=int, default=5_000_000)
    ap.add_argument("--object", type=int, default=200_000)
    ap.add_argument("--sh-degree", type=int, default=0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        RL.bind_to_gpu_numa_node(local)
    W, H = 1920, 1080
    g_scene = scene.synthetic_gaussians(args.scene, seed=1234, extent=(4, 4, 1), log_scale_mean=math.log(0.006), log_scale_std=0.5,
                                        opacity_mean=0.0, opacity_std=2.0, sh_degree=4)
    g_obj = scene.synthetic_gaussians(args.object, seed=77, extent=(0.4, 0.4, 0.4), log_scale_mean=math.log(0.004), log_scale_std=0.4,
                                      opacity_mean=1.0, opacity_std=1.0, sh_degree=4)
    raw_scene, M = to_raw(g_scene)
    raw_obj, _ = to_raw(g_obj)
    t0 = time.perf_counter()
    rs = edit.ResidentScene(raw_scene, {"obj": raw_obj}, dev)
    torch.cuda.synchronize()
    t_load = time.perf_counter() - t0
    del g_scene, raw_scene
    cams_all = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=300))
    idx = RL.shard_indices(min(300, args.frames * world), rank, world)
    cams = RL.pack_cameras([cams_all[i] for i in idx])

    def transform(i):  # the object circles the origin, spinning about z, growing slightly
        a = 2 * math.pi * idx[i] / 300.0
        c, s = math.cos(3 * a), math.sin(3 * a)
        R = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        return {"obj": (torch.tensor([1.5 * math.cos(a), 1.5 * math.sin(a), 0.2]), R, 1.0 + 0.3 * math.sin(a), torch.zeros(3))}

    loop = RL.FrameLoop(rs.compose(transform(0)), args.sh_degree, W, H, device=dev, ring=3, to_host=True, product=True, pack8=True)
    loop.render(cams[:6], before_frame=lambda i: rs.compose(transform(i)))  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    chk = [0]
    t0 = time.perf_counter()
    stats = loop.render(cams, lambda i, fr, st: chk.__setitem__(0, chk[0] + int(fr["rgba8"][H // 2, W // 2, 0])), before_frame=lambda i: rs.compose(transform(i)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    # device-only time of the per-frame edit (compose) for the record
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        rs.compose(transform(i % len(idx)))
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"metric": "end-to-end product frames/s, config 5 stand-in", "value": len(idx) * world / dt, "unit": "frames/s", "n_gpus": world,
                          "frames_per_rank": len(idx), "config": {"scene_gaussians": args.scene, "object_gaussians": args.object, "M": M,
                                                                    "sh_degree_rendered": args.sh_degree, "resolution": [W, H]},
                          "per_frame": "ResidentScene.compose (gsr_activate_gaussians on the object) + render() (axis normals, 6-channel forward, normal maps) "
                                       "+ 8-bit pack + D2H of %d bytes" % loop.d2h_bytes_per_frame,
                          "compose_ms": e0.elapsed_time(e1) / 20, "scene_activation_s": t_load,
                          "avg_num_rendered": sum(s["num_rendered"] for s in stats) / len(stats), "rerendered": loop.rerendered, "data": "synthetic"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

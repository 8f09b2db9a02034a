"""Times forward(for_backward) and backward of the 3M-Gaussian 1080p workload with CUDA events (device time only)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovfx_b200 import scene  # noqa: E402
from tests import helpers as Hh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--gaussians", type=int, default=3_000_000)
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = scene.config3_scene(P=args.gaussians)
    cams = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=300))
    fw_ms, bw_ms = [], []
    for i in range(args.iters + 3):
        a = Hh.resolve(dict(g=g, cam=cams[(i * 7) % 300], sh_degree=3, bg=(0.0, 0.0, 0.0), scale_modifier=1.0), dev)
        dc, dd, da = Hh.image_grads(a, device=dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if args.reference:
            from oracle import ref_cuda
            e[0].record()
            fw = Hh.run_ref(a)
            e[1].record()
            ref_cuda.backward(fw, dc, dd, da)
            e[2].record()
        else:
            from autovfx_b200.rasterizer import GaussianRasterizer
            leaves = {k: a[k].detach().clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
            m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
            rast = GaussianRasterizer(Hh.settings_from(a))
            e[0].record()
            color, depth, alpha, radii = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
            e[1].record()
            torch.autograd.backward([color, depth, alpha], [dc, dd, da])
            e[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            fw_ms.append(e[0].elapsed_time(e[1]))
            bw_ms.append(e[1].elapsed_time(e[2]))
    fw_ms.sort()
    bw_ms.sort()
    print("%s: forward(for_backward) median %.3f ms, backward median %.3f ms, sum %.3f ms" %
          ("reference" if args.reference else "ours", fw_ms[len(fw_ms) // 2], bw_ms[len(bw_ms) // 2], fw_ms[len(fw_ms) // 2] + bw_ms[len(bw_ms) // 2]))


if __name__ == "__main__":
    main()

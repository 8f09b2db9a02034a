"""Renders a few frames of the 3M-Gaussian 1080p workload (for ncu captures; no timing is taken here)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovfx_b200 import scene  # noqa: E402
from tests import helpers as Hh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--gaussians", type=int, default=3_000_000)
    ap.add_argument("--backward", action="store_true")
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--product", action="store_true", help="the fused product frame: axis_normals + 6-channel forward + normal maps")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = scene.config3_scene(P=args.gaussians)
    cams = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=300))
    for i in range(args.frames):
        a = Hh.resolve(dict(g=g, cam=cams[i * 7], sh_degree=3, bg=(0.0, 0.0, 0.0), scale_modifier=1.0), dev)
        if args.product:
            from autovfx_b200 import rasterizer as R, renderer as RD
            nrm = RD.axis_normals(a["means3D"], a["scales"], a["rotations"], a["campos"], remap01=True)
            res = R.forward_multi(a["means3D"], a["shs"], None, nrm, a["opacities"], a["scales"], a["rotations"], None, Hh.settings_from(a), sync=True)
            RD.normal_maps(res[3], res[1][0], torch.linalg.inv_ex(a["view"])[0], a["W"] / (2 * a["tanfovx"]), a["H"] / (2 * a["tanfovy"]),
                           a["W"] / 2, a["H"] / 2)
            print("product frame", i, res[5].stats())
        elif args.reference:
            Hh.run_ref(a)
        elif args.backward:
            dc, dd, da = Hh.image_grads(a, device=dev)
            Hh.ours_backward(a, dc, dd, da)
        else:
            o = Hh.run_ours(a, debug=False)
            print("frame", i, o["stats"])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()

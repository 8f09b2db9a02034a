"""Generates the golden vectors under tests/golden/ by running the reference's OWN CUDA rasterizer
(oracle/_ref/libref_dgr.so = unmodified diff-gaussian-rasterization compiled from /root/reference) on a B200.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'     # then copy gpurun_out/golden/*.npz here

Inputs are the deterministic named cases of tests/helpers.py (seeded torch generators), so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import helpers as Hh  # noqa: E402
from oracle import ref_cuda  # noqa: E402

CASES = {  # name -> store gradients?
    "config1": ("fw+small_grads",), "small_sh": ("fw+grads",), "small_deg1_m25": ("fw+grads",), "small_precomp": ("fw+grads",),
    "big_splats": ("fw+grads",), "dense_tile": ("fw",), "coplanar": ("fw",),
}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, (mode,) in CASES.items():
        a = Hh.resolve(Hh.case_inputs(name), dev)
        fw = ref_cuda.forward(a["means3D"], a["opacities"], a["view"], a["proj"], a["campos"], a["W"], a["H"], a["tanfovx"], a["tanfovy"],
                              shs=a["shs"], colors_precomp=a["colors_precomp"], scales=a["scales"], rotations=a["rotations"],
                              cov3D_precomp=a["cov3D_precomp"], sh_degree=a["sh_degree"], scale_modifier=a["scale_modifier"], bg=a["bg"])
        st = ref_cuda.state(dev)
        torch.cuda.synchronize()
        out = {"case": np.array(name), "radii": fw["radii"].cpu().numpy(), "num_rendered": np.array(fw["num_rendered"]),
               "color": fw["color"].cpu().numpy(), "depth": fw["depth"].cpu().numpy(), "alpha": fw["alpha"].cpu().numpy(),
               "point_list": st["point_list"].cpu().numpy(), "ranges": st["ranges"].cpu().numpy(), "n_contrib": st["n_contrib"].cpu().numpy(),
               "means2D": st["means2D"].cpu().numpy(), "depths": st["depths"].cpu().numpy(), "conic_opacity": st["conic_opacity"].cpu().numpy()}
        if mode != "fw":
            dc, dd, da = Hh.image_grads(a, device=dev)
            g = ref_cuda.backward(fw, dc, dd, da)
            torch.cuda.synchronize()
            keys = ["dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dcolors", "dL_dcov3D"]
            if mode == "fw+grads":
                keys.append("dL_dsh")
            for k in keys:
                out[k] = g[k].cpu().numpy()
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, "%.0f KB" % (os.path.getsize(path) / 1024), "R =", fw["num_rendered"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))

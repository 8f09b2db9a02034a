"""GPU diagnostic: product CUDA path vs the compiled reference (oracle/_ref) vs the CPU oracle, stage by stage.
Prints a report (never stops at the first mismatch).  Run on the GPU box:

    python tools/parity_report.py [--big] > gpurun_out/parity_report.txt
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as Hh  # noqa: E402
from autovfx_b200 import scene  # noqa: E402


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> int:
    """number of elements whose fp32 bit patterns differ"""
    return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())


def compare_case(name, a, do_oracle=True, do_backward=True):
    dev = a["means3D"].device
    P = a["means3D"].shape[0]
    print("\n=== case %s: P=%d %dx%d D=%d ===" % (name, P, a["W"], a["H"], a["sh_degree"]))
    ours = Hh.run_ours(a, for_backward=True, sorted_keys=True)
    torch.cuda.synchronize()
    ref = Hh.run_ref(a)
    from oracle import ref_cuda
    rs = ref_cuda.state(dev)
    torch.cuda.synchronize()
    v = ours["views"]
    vis_r = ref["radii"] > 0
    vis_o = ours["radii"] > 0
    print("stats ours:", ours["stats"], " ref R:", ref["num_rendered"], " P_vis ref:", int(vis_r.sum()))
    print("radii mismatches: %d / %d" % (int((ours["radii"] != ref["radii"]).sum()), P))
    both = vis_r & vis_o
    rec = v["records"]
    for nm, mine, theirs in (("means2D.x", rec[:, 0], rs["means2D"][:, 0]), ("means2D.y", rec[:, 1], rs["means2D"][:, 1]),
                             ("conic.a", rec[:, 2], rs["conic_opacity"][:, 0]), ("conic.b", rec[:, 3], rs["conic_opacity"][:, 1]),
                             ("conic.c", rec[:, 4], rs["conic_opacity"][:, 2]), ("opacity", rec[:, 5], rs["conic_opacity"][:, 3]),
                             ("depth", rec[:, 6], rs["depths"])):
        print("  %-10s bit-diffs %d   maxabs %.3e" % (nm, bits_equal(mine[both], theirs[both]), Hh.maxabs(mine[both], theirs[both])))
    if a["colors_precomp"] is None:
        for c in range(3):
            print("  rgb[%d]     bit-diffs %d   maxabs %.3e" % (c, bits_equal(rec[:, 8 + c][both], rs["rgb"][:, c][both]),
                                                                Hh.maxabs(rec[:, 8 + c][both], rs["rgb"][:, c][both])))
        cl = torch.stack([(v["clamped"] >> c) & 1 for c in range(3)], dim=1)
        print("  clamped mismatches:", int((cl[both] != rs["clamped"][both]).sum()))
    if a["cov3D_precomp"] is None:
        print("  cov3D      bit-diffs %d   maxabs %.3e" % (bits_equal(v["cov3D"][both], rs["cov3D"][both]), Hh.maxabs(v["cov3D"][both], rs["cov3D"][both])))
    R = ref["num_rendered"]
    print("R equal:", ours["stats"]["num_rendered"] == R)
    if ours["stats"]["num_rendered"] == R:
        pl_o = v["point_list"][:R]
        print("  point_list mismatches: %d / %d" % (int((pl_o != rs["point_list"]).sum()), R))
        print("  ranges mismatches: %d" % int((v["ranges"] != rs["ranges"]).sum()))
        # rebuild the reference key (tile << 32 | depth bits) from our (depth bits << 32 | id) + ranges
        sk = v["sorted_keys"][:R]
        depth_bits = (sk >> 32) & 0xFFFFFFFF
        rg = v["ranges"].long()
        tile_of = torch.zeros(R, dtype=torch.int64, device=dev)
        cnt = (rg[:, 1] - rg[:, 0])
        tile_ids = torch.repeat_interleave(torch.arange(rg.shape[0], device=dev), cnt)
        starts = torch.repeat_interleave(rg[:, 0], cnt)
        order = torch.argsort(starts, stable=True)
        tile_of = tile_ids[order] if R else tile_of
        key_o = (tile_of << 32) | depth_bits
        print("  sorted key mismatches: %d" % int((key_o != rs["point_list_keys"]).sum()))
        print("  n_contrib mismatches: %d / %d" % (int((v["n_contrib"] != rs["n_contrib"]).sum()), a["W"] * a["H"]))
    for nm in ("color", "depth", "alpha"):
        print("  image %-6s maxabs %.3e  bit-diffs %d / %d" % (nm, Hh.maxabs(ours[nm], ref[nm]), bits_equal(ours[nm], ref[nm]), ours[nm].numel()))
    if do_oracle:
        t = time.time()
        orc = Hh.run_oracle(a)
        print("oracle (CPU) %.2fs: radii mismatches vs ref %d, R %d vs %d" % (time.time() - t, int((torch.from_numpy(orc["radii"]).to(dev) != ref["radii"]).sum()),
                                                                              orc["num_rendered"], R))
        for nm in ("color", "depth", "alpha"):
            print("  oracle image %-6s maxabs vs ref %.3e   vs ours %.3e" % (nm, Hh.maxabs(orc[nm], ref[nm]), Hh.maxabs(orc[nm], ours[nm])))
    if do_backward:
        dc, dd, da = Hh.image_grads(a, device=dev)
        gr = ref_cuda.backward(ref, dc, dd, da)
        torch.cuda.synchronize()
        _, go = Hh.ours_backward(a, dc, dd, da)
        torch.cuda.synchronize()
        pairs = [("means3D", go["means3D"], gr["dL_dmeans3D"]), ("means2D", go["means2D"], gr["dL_dmeans2D"]), ("opacities", go["opacities"], gr["dL_dopacity"])]
        if a["shs"] is not None:
            pairs.append(("shs", go["shs"], gr["dL_dsh"]))
        else:
            pairs.append(("colors_precomp", go["colors_precomp"], gr["dL_dcolors"]))
        if a["scales"] is not None:
            pairs += [("scales", go["scales"], gr["dL_dscales"]), ("rotations", go["rotations"], gr["dL_drotations"])]
        else:
            pairs.append(("cov3D_precomp", go["cov3D_precomp"], gr["dL_dcov3D"]))
        for nm, x, y in pairs:
            print("  grad %-14s rel-to-max err %.3e   maxabs %.3e  (|ref|max %.3e)" % (nm, Hh.relerr(x, y), Hh.maxabs(x, y), float(y.abs().max())))
        if do_oracle:
            og = Hh.oracle_backward(a, orc, dc, dd, da)
            print("  oracle grads vs ref: " + ", ".join("%s %.2e" % (k, Hh.relerr(og[k], gr[k])) for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dsh", "dL_dcov3D") if gr[k].numel()))


def time_both(a, iters=20):
    from oracle import ref_cuda
    from autovfx_b200 import rasterizer as R
    s = Hh.settings_from(a)
    def ours():
        return R.forward_raw(a["means3D"], a["shs"], a["colors_precomp"], a["opacities"], a["scales"], a["rotations"], a["cov3D_precomp"], s, sync=False)
    def ref():
        return Hh.run_ref(a)
    for nm, fn in (("ours", ours), ("ref", ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("  timing %-5s %.3f ms/frame" % (nm, e0.elapsed_time(e1) / iters))
    print("  ours stats:", R.last_frame_stats())
    # forward + backward (BASELINE config 3): public autograd API vs the reference's forward+backward
    dc, dd, da = Hh.image_grads(a, device=a["means3D"].device)
    def ours_fb():
        Hh.ours_backward(a, dc, dd, da)
    def ref_fb():
        ref_cuda.backward(Hh.run_ref(a), dc, dd, da)
    for nm, fn in (("ours fwd+bwd", ours_fb), ("ref  fwd+bwd", ref_fb)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("  timing %-13s %.3f ms/iter" % (nm, e0.elapsed_time(e1) / 5))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--cases", default="config1,small_sh,small_deg1_m25,small_precomp,big_splats,dense_tile,coplanar")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    print("device:", torch.cuda.get_device_name(0))
    for name in [c for c in args.cases.split(",") if c]:
        try:
            compare_case(name, Hh.resolve(Hh.case_inputs(name), dev))
        except Exception:  # noqa: BLE001
            print("!! case %s raised:" % name)
            traceback.print_exc(file=sys.stdout)
    if args.big:
        try:
            t = time.time()
            g = scene.config3_scene()
            cams = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=300))
            print("\nscene gen %.1fs" % (time.time() - t))
            for ci in (0, 77, 150):
                case = dict(g=g, cam=cams[ci], sh_degree=3, bg=(0.0, 0.0, 0.0), scale_modifier=1.0)
                a = Hh.resolve(case, dev)
                compare_case("3M_cam%d" % ci, a, do_oracle=False, do_backward=(ci == 0))
                time_both(a)
        except Exception:  # noqa: BLE001
            print("!! big case raised:")
            traceback.print_exc(file=sys.stdout)


if __name__ == "__main__":
    main()

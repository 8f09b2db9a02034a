"""TEST INFRASTRUCTURE — torch/ctypes front end of the compiled, UNMODIFIED reference rasterizer
(oracle/_ref/libref_dgr.so, built by oracle/Makefile from the sources under /root/reference).

Used by the GPU parity tests, tests/golden/make_golden.py and bench.py's reference arm.  Never imported by
the product package.  All launches go to the legacy default stream like the reference; callers must be on
torch's default stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_dgr.so")
_lib = None


class RefState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("depths", "clamped", "internal_radii", "means2D", "cov3D", "conic_opacity", "rgb",
                                          "point_offsets", "tiles_touched", "point_list", "point_list_keys", "ranges", "n_contrib")]
    _fields_ += [("P", C.c_int), ("R", C.c_int), ("W", C.c_int), ("H", C.c_int)]


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_dgr.so missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(SO)
        _lib.ref_forward.restype = C.c_int
        _lib.ref_backward.restype = C.c_int
        _lib.ref_last_error.restype = C.c_char_p
    return _lib


def _p(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _f(t):
    return None if t is None else t.detach().float().contiguous()


def forward(means3D, opacities, view, proj, campos, W, H, tanfovx, tanfovy, *, shs=None, colors_precomp=None, scales=None,
            rotations=None, cov3D_precomp=None, sh_degree=3, scale_modifier=1.0, bg=None, prefiltered=False, debug=False):
    L = lib()
    dev = means3D.device
    means3D, opacities = _f(means3D), _f(opacities)
    shs, colors_precomp, scales, rotations, cov3D_precomp = _f(shs), _f(colors_precomp), _f(scales), _f(rotations), _f(cov3D_precomp)
    view, proj, campos = _f(view), _f(proj), _f(campos)
    bg = torch.zeros(3, device=dev) if bg is None else _f(bg)
    P = means3D.size(0)
    M = 0 if shs is None else shs.size(1)
    color = torch.empty((3, H, W), device=dev)
    depth = torch.empty((1, H, W), device=dev)
    alpha = torch.empty((1, H, W), device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    R = L.ref_forward(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(bg), C.c_int(W), C.c_int(H), _p(means3D), _p(shs),
                      _p(colors_precomp), _p(opacities), _p(scales), C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp),
                      _p(view), _p(proj), _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), C.c_int(int(prefiltered)),
                      _p(color), _p(depth), _p(alpha), _p(radii) if P else None, C.c_int(int(debug)))
    if R < 0:
        raise RuntimeError("ref_forward: " + L.ref_last_error().decode())
    keep = dict(means3D=means3D, opacities=opacities, shs=shs, colors_precomp=colors_precomp, scales=scales, rotations=rotations,
                cov3D_precomp=cov3D_precomp, view=view, proj=proj, campos=campos, bg=bg)
    return {"color": color, "depth": depth, "alpha": alpha, "radii": radii, "num_rendered": R, "_keep": keep,
            "_cfg": (P, sh_degree, M, W, H, tanfovx, tanfovy, scale_modifier)}


def state(device) -> Dict[str, torch.Tensor]:
    """Copies of the reference's per-stage buffers after the last forward (cloned: the buffers are reused)."""
    s = RefState()
    lib().ref_get_state(C.byref(s))
    P, R, W, H = s.P, s.R, s.W, s.H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)

    def grab(ptr, n, dtype):
        if n == 0 or not ptr:
            return torch.empty(0, dtype=dtype, device=device)
        out = torch.empty(n, dtype=dtype, device=device)
        rc = lib().ref_copy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), C.c_size_t(out.numel() * out.element_size()))
        if rc != 0:
            raise RuntimeError("ref_copy failed: cuda error %d" % rc)
        return out

    return {
        "depths": grab(s.depths, P, torch.float32), "clamped": grab(s.clamped, 3 * P, torch.uint8).view(P, 3),
        "means2D": grab(s.means2D, 2 * P, torch.float32).view(P, 2), "cov3D": grab(s.cov3D, 6 * P, torch.float32).view(P, 6),
        "conic_opacity": grab(s.conic_opacity, 4 * P, torch.float32).view(P, 4), "rgb": grab(s.rgb, 3 * P, torch.float32).view(P, 3),
        "point_offsets": grab(s.point_offsets, P, torch.int32), "tiles_touched": grab(s.tiles_touched, P, torch.int32),
        "point_list": grab(s.point_list, R, torch.int32), "point_list_keys": grab(s.point_list_keys, R, torch.int64),
        "ranges": grab(s.ranges, 2 * tiles, torch.int32).view(tiles, 2), "n_contrib": grab(s.n_contrib, W * H, torch.int32).view(H, W),
    }


def backward(fw, dL_dcolor, dL_ddepth, dL_dalpha, debug=False) -> Dict[str, torch.Tensor]:
    L = lib()
    k = fw["_keep"]
    P, D, M, W, H, tanx, tany, mod = fw["_cfg"]
    dev = k["means3D"].device
    z = lambda *shape: torch.empty(shape, device=dev)  # noqa: E731
    g = {"dL_dmeans2D": z(P, 3), "dL_dconic": z(P, 2, 2), "dL_dopacity": z(P, 1), "dL_dcolors": z(P, 3), "dL_ddepths": z(P, 1),
         "dL_dmeans3D": z(P, 3), "dL_dcov3D": z(P, 6), "dL_dsh": z(P, M, 3), "dL_dscales": z(P, 3), "dL_drotations": z(P, 4)}
    rc = L.ref_backward(C.c_int(P), C.c_int(D), C.c_int(M), _p(k["bg"]), C.c_int(W), C.c_int(H), _p(k["means3D"]), _p(k["shs"]),
                        _p(k["colors_precomp"]), _p(k["scales"]), C.c_float(mod), _p(k["rotations"]), _p(k["cov3D_precomp"]),
                        _p(k["view"]), _p(k["proj"]), _p(k["campos"]), C.c_float(tanx), C.c_float(tany), _p(fw["radii"]),
                        _p(fw["alpha"]), _p(_f(dL_dcolor)), _p(_f(dL_ddepth)), _p(_f(dL_dalpha)),
                        _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_ddepths"]),
                        _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]) if M else None, _p(g["dL_dscales"]),
                        _p(g["dL_drotations"]), C.c_int(int(debug)))
    if rc != 0:
        raise RuntimeError("ref_backward: " + L.ref_last_error().decode())
    return g


def mark_visible(means3D, view, proj):
    P = means3D.size(0)
    out = torch.empty((P,), dtype=torch.uint8, device=means3D.device)
    lib().ref_mark_visible(C.c_int(P), _p(_f(means3D)), _p(_f(view)), _p(_f(proj)), _p(out))
    return out.bool()


def dist2(points):
    P = points.size(0)
    out = torch.empty((P,), device=points.device)
    rc = lib().ref_dist2(C.c_int(P), _p(_f(points)), _p(out))
    if rc != 0:
        raise RuntimeError("ref_dist2: " + lib().ref_last_error().decode())
    return out

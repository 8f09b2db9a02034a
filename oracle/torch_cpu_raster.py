"""TEST / BASELINE INFRASTRUCTURE — a pure-CPU PyTorch rasterize loop (BASELINE.md "R3", north_star's "pure-CPU PyTorch rasterize
loop timed on the box's host cores").  Vectorised torch ops on CPU tensors, float32: preprocess for all Gaussians at once
(forward.cu:155-256), tile membership from the reference's rectangle rule (auxiliary.h:46-56), per-tile depth sort, and the blend
recurrence (forward.cu:330-366) evaluated per tile as [pixels x splats] matrices with an exclusive cumulative product for the
transmittance and the three skip / terminate rules as masks.  It is a baseline to be timed, checked against the C oracle to 1e-4
on small scenes (tests/test_oracle_cpu.py); it is never used by the product."""
from __future__ import annotations

import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435]


def _sh(deg, sh, d):
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6] + SH_C2[3] * xz * sh[:, 7] + \
            SH_C2[4] * (xx - yy) * sh[:, 8]
    if deg > 2:
        res = res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10] + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + \
            SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + \
            SH_C3[5] * z * (xx - yy) * sh[:, 14] + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15]
    return res


@torch.no_grad()
def rasterize(means3D, scales, rotations, opacities, shs, view, proj, campos, W, H, tanfovx, tanfovy, sh_degree=3, scale_modifier=1.0,
              bg=(0.0, 0.0, 0.0)):
    """Returns color [3,H,W], depth [1,H,W], alpha [1,H,W], radii [P] (int32).  All inputs CPU float32 tensors."""
    P = means3D.shape[0]
    hom = torch.cat([means3D, torch.ones(P, 1)], dim=1)
    p_hom = hom @ proj
    p_w = 1.0 / (p_hom[:, 3] + 0.0000001)
    p_proj = p_hom[:, :3] * p_w[:, None]
    t = (hom @ view)[:, :3]
    tz = t[:, 2]
    front = tz > 0.2
    r, x, y, z = rotations.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).view(P, 3, 3)
    L = R * (scales * scale_modifier).unsqueeze(1)
    Sigma = L @ L.transpose(1, 2)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    tzs = torch.where(front, tz, torch.ones_like(tz))
    tx = torch.clamp(t[:, 0] / tzs, -1.3 * tanfovx, 1.3 * tanfovx) * tzs
    ty = torch.clamp(t[:, 1] / tzs, -1.3 * tanfovy, 1.3 * tanfovy) * tzs
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tzs, zero, -(fx * tx) / (tzs * tzs), zero, fy / tzs, -(fy * ty) / (tzs * tzs)], dim=-1).view(P, 2, 3)
    T = J @ view[:3, :3].T
    cov = T @ Sigma @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    ok = front & (det != 0)
    dinv = 1.0 / torch.where(det != 0, det, torch.ones_like(det))
    ca, cb, cc = c * dinv, -b * dinv, a * dinv
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam))
    px = (((p_proj[:, 0].double() + 1.0) * W - 1.0) * 0.5).float()
    py = (((p_proj[:, 1].double() + 1.0) * H - 1.0) * 0.5).float()
    gx, gy = (W + 15) // 16, (H + 15) // 16
    x0 = torch.clamp(((px - radius) / 16).to(torch.int32), 0, gx)
    y0 = torch.clamp(((py - radius) / 16).to(torch.int32), 0, gy)
    x1 = torch.clamp(((px + radius + 15) / 16).to(torch.int32), 0, gx)
    y1 = torch.clamp(((py + radius + 15) / 16).to(torch.int32), 0, gy)
    vis = ok & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(vis, radius, torch.zeros_like(radius)).to(torch.int32)
    d = means3D - campos
    d = d / d.norm(dim=1, keepdim=True)
    rgb = torch.clamp_min(_sh(sh_degree, shs, d) + 0.5, 0.0)
    op = opacities.reshape(-1)
    bgt = torch.tensor(bg, dtype=torch.float32)
    color = torch.zeros(3, H, W)
    depth = torch.zeros(1, H, W)
    alpha = torch.zeros(1, H, W)
    idx_vis = torch.nonzero(vis).squeeze(1)
    order = idx_vis[torch.argsort(tz[idx_vis], stable=True)]  # global front-to-back order restricted per tile below
    ox0, oy0, ox1, oy1 = x0[order], y0[order], x1[order], y1[order]
    ys, xs = torch.meshgrid(torch.arange(16, dtype=torch.float32), torch.arange(16, dtype=torch.float32), indexing="ij")
    for tyi in range(gy):
        row = (oy0 <= tyi) & (oy1 > tyi)
        cand = torch.nonzero(row).squeeze(1)
        if cand.numel() == 0:
            color[:, tyi * 16:(tyi + 1) * 16] = bgt.view(3, 1, 1)
            continue
        cx0, cx1 = ox0[cand], ox1[cand]
        for txi in range(gx):
            sel = cand[(cx0 <= txi) & (cx1 > txi)]
            h, w = min(16, H - tyi * 16), min(16, W - txi * 16)
            if sel.numel() == 0:
                color[:, tyi * 16:tyi * 16 + h, txi * 16:txi * 16 + w] = bgt.view(3, 1, 1)
                continue
            g = order[sel]
            pxs = (xs[:h, :w] + txi * 16).reshape(-1, 1)
            pys = (ys[:h, :w] + tyi * 16).reshape(-1, 1)
            dx, dy = px[g][None, :] - pxs, py[g][None, :] - pys
            power = -0.5 * (ca[g][None, :] * dx * dx + cc[g][None, :] * dy * dy) - cb[g][None, :] * dx * dy
            al = torch.clamp_max(op[g][None, :] * torch.exp(power), 0.99)
            hit = (power <= 0) & (al >= 1.0 / 255.0)
            al = torch.where(hit, al, torch.zeros_like(al))
            Tafter = torch.cumprod(1 - al, dim=1)  # transmittance after each splat if nothing terminated
            dead = (hit & (Tafter < 0.0001)).to(torch.int32).cumsum(1) > 0  # from the first splat whose blend would drop T below 1e-4
            al = torch.where(dead, torch.zeros_like(al), al)
            Tbefore = torch.cat([torch.ones(al.shape[0], 1), torch.cumprod(1 - al, dim=1)[:, :-1]], dim=1)
            wgt = al * Tbefore
            Tfin = torch.prod(1 - al, dim=1)
            col = wgt @ rgb[g] + Tfin[:, None] * bgt[None, :]
            color[:, tyi * 16:tyi * 16 + h, txi * 16:txi * 16 + w] = col.T.reshape(3, h, w)
            depth[0, tyi * 16:tyi * 16 + h, txi * 16:txi * 16 + w] = (wgt @ tz[g]).reshape(h, w)
            alpha[0, tyi * 16:tyi * 16 + h, txi * 16:txi * 16 + w] = (1 - Tfin).reshape(h, w)
    return color, depth, alpha, radii

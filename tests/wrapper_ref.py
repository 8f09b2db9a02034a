"""Torch restatement, op for op, of the reference's render() wrapper helpers — test infrastructure.

The originals cannot be imported (gaussian_renderer/__init__.py imports kornia; utils/general_utils.py:83 hard-codes
device='cuda'), so they are restated here with a ``device`` that follows the inputs.  Executed with torch's CUDA kernels on
the GPU box they are the closest available stand-in for "the reference run here" for this row; on CPU they pin
oracle/render_oracle.py.  "GR/" = sugar/gaussian_splatting/gaussian_renderer/__init__.py, "GU/" = .../utils/general_utils.py.
"""
from __future__ import annotations

import math

import torch


def build_rotation(quats):
    """Rotation matrices [N,3,3] of quaternions (real part first) that are normalised first — the semantics and the per-element
    operation order of GU/:78-99 (each entry is formed by the same products, sums and the final `1 - 2*(..)` / `2*(..)`)."""
    n = torch.sqrt(quats[:, 0] * quats[:, 0] + quats[:, 1] * quats[:, 1] + quats[:, 2] * quats[:, 2] + quats[:, 3] * quats[:, 3])
    w, x, y, z = (quats / n[:, None]).unbind(-1)
    rows = [torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)), -1),
            torch.stack((2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)), -1),
            torch.stack((2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), -1)]
    return torch.stack(rows, 1)


def get_minimum_axis(scales, rotations):
    """Column of the rotation matrix that belongs to the smallest scale (GU/:136-141: argsort ascending, gather, first column)."""
    order = torch.argsort(scales, dim=-1, descending=False)
    cols = torch.gather(build_rotation(rotations), 2, order[:, None, :].expand(-1, 3, -1))
    return cols[:, :, 0]


def flip_align_view(normal, viewdir):
    """Flip the normals that point away from the viewer (GU/:151-157): sign of sum(normal * -viewdir)."""
    facing = (normal * -viewdir).sum(dim=-1, keepdim=True) >= 0
    return normal * torch.where(facing, 1, -1), facing


def get_normal(xyz, scales, rotations, campos):  # scene/gaussian_model.py:120-124 + GR/:131-132
    dir_pp = xyz - campos.repeat(xyz.shape[0], 1)
    dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    normal_axis = get_minimum_axis(scales, rotations)
    normal_axis, _ = flip_align_view(normal_axis, dir_pp_normalized)
    return normal_axis / normal_axis.norm(dim=1, keepdim=True)


def fov2focal(fov, pixels):  # utils/graphics_utils.py:74-75
    return pixels / (2 * math.tan(fov / 2))


def get_ray_directions(H, W, K, device):  # GR/:41-80, create_meshgrid(H, W, False): x = column, y = row, pixel units
    xs = torch.linspace(0, W - 1, W, device=device, dtype=torch.float32)
    ys = torch.linspace(0, H - 1, H, device=device, dtype=torch.float32)
    v, u = torch.meshgrid(ys, xs, indexing="ij")
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    return torch.stack([(u - cx + 0.5) / fx, (v - cy + 0.5) / fy, torch.ones_like(u)], -1)


def depth_pcd2normal(points):
    """Pseudo normals of an [H,W,3] point map: normalised cross product of the horizontal and vertical central differences,
    zero on the one-pixel border (the computation of GR/:23-38)."""
    H, W, _ = points.shape
    horiz = points[1:H - 1, 2:W] - points[1:H - 1, 0:W - 2]     # right - left
    vert = points[0:H - 2, 1:W - 1] - points[2:H, 1:W - 1]      # top - bottom
    n = torch.nn.functional.normalize(torch.cross(horiz, vert, dim=-1), p=2, dim=-1)
    out = torch.zeros_like(points)
    if H > 2 and W > 2:
        out[1:H - 1, 1:W - 1] = n
    return out


def normal_image(normal_img_chw):  # GR/:168-176
    n = (normal_img_chw - 0.5) * 2.
    return torch.nn.functional.normalize(n.permute(1, 2, 0), p=2, dim=-1)


def pseudo_normal(depth_hw, world_view_transform, FoVx, FoVy):  # GR/:178-191
    h, w = depth_hw.shape
    fx, fy = fov2focal(FoVx, w), fov2focal(FoVy, h)
    cx, cy = w / 2, h / 2
    c2w = world_view_transform.inverse()
    directions = get_ray_directions(h, w, torch.FloatTensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]]), depth_hw.device)
    rays_d = directions @ c2w[:3, :3].T
    rays_o = c2w[:3, 3].expand_as(rays_d)
    points3D = rays_o + rays_d * depth_hw.unsqueeze(-1)
    return depth_pcd2normal(points3D)


def save_image_bytes(img_chw):  # torchvision.utils.save_image for one image: mul(255).add_(0.5).clamp_(0,255).permute(1,2,0).to(uint8)
    return img_chw.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)


def render_two_pass(rasterize, xyz, shs, opacity, scales, rotations, sh_degree, cam, bg):
    """GR/:83-218 with ``rasterize(shs=..., colors_precomp=...) -> (color, depth, alpha, radii)`` standing for the
    reference's GaussianRasterizer call (tests bind it to the compiled reference rasterizer)."""
    campos = cam["campos"]
    normal = get_normal(xyz, scales, rotations, campos)
    normal_normed = normal * 0.5 + 0.5
    rendered_image, depth_image, alpha_image, radii = rasterize(shs=shs, colors_precomp=None)
    rendered_image = torch.cat((rendered_image, alpha_image), dim=0)
    depth_image = depth_image.squeeze(0)
    nimg = rasterize(shs=None, colors_precomp=normal_normed)[0]
    return {"render": rendered_image, "depth": depth_image, "normal": normal_image(nimg),
            "pseudo_normal": pseudo_normal(depth_image, cam["viewmatrix"], cam["FoVx"], cam["FoVy"]),
            "normal_normed": normal_normed, "normal_raw_image": nimg, "radii": radii}


# ----------------------------------------------------------------------------- per-frame edit path (gaussians_utils.py, rotation_utils.py)
def matrix_to_quaternion(matrix):
    """(w,x,y,z) with w >= 0 of a rotation matrix via scipy (an independent implementation; rotation_utils.py:24-84 returns the
    same quaternion up to the sign convention and float32 round-off)."""
    from scipy.spatial.transform import Rotation
    x, y, z, w = Rotation.from_matrix(matrix.detach().cpu().double().numpy()).as_quat()
    q = torch.tensor([w, x, y, z], dtype=torch.float64)
    return (q if w >= 0 else -q).to(dtype=matrix.dtype, device=matrix.device)


def quaternion_multiply(a, b):
    """Hamilton product a*b (real part first), result flipped to a non-negative real part — the semantics of
    rotation_utils.py:113-150, with the same left-to-right evaluation of each component."""
    a0, a1, a2, a3 = a.unbind(-1)
    b0, b1, b2, b3 = b.unbind(-1)
    prod = torch.stack((a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
                        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0), -1)
    return torch.where(prod[..., :1] < 0, -prod, prod)


def transform_gaussians(raw, center, rotation, scaling, initial_center, quat=None):
    """The op sequence of gaussians_utils.py:88-125 on a dict of raw tensors: scale about the pivot (and shift the log-scales),
    rotate about the pivot (matmul with R^T, quaternion product), translate by (center - pivot)."""
    import numpy as np
    pivot = initial_center.unsqueeze(0)
    p = raw["xyz"].clone()
    p -= pivot
    p *= scaling
    p += pivot
    log_s = raw["scaling"].clone()
    log_s += np.log(scaling)
    p -= pivot
    p = torch.matmul(p, rotation.T)
    p += pivot
    q = quaternion_multiply(matrix_to_quaternion(rotation) if quat is None else quat, raw["rotation"].clone())
    p += (center - initial_center).unsqueeze(0)
    out = dict(raw)
    out["xyz"], out["rotation"], out["scaling"] = p, q, log_s
    return out


def merge_two_gaussians(r1, r2):  # gaussians_utils.py:71-84
    return {k: torch.cat([r1[k], r2[k]], dim=0) for k in ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")}


def activate(raw):  # scene/gaussian_model.py:95-115: get_xyz, get_features, get_opacity, get_scaling, get_rotation
    return {"means3D": raw["xyz"], "shs": torch.cat((raw["f_dc"], raw["f_rest"]), dim=1), "opacities": torch.sigmoid(raw["opacity"]),
            "scales": torch.exp(raw["scaling"]), "rotations": torch.nn.functional.normalize(raw["rotation"])}

"""Synthetic scenes, cameras, trajectories and the 3DGS ``.ply`` layout.

Host-side helpers that reproduce the CALLER contracts of the reference hot path so that benches and
parity tests feed the rasterizer exactly what AutoVFX feeds it (SURVEY §8a "camera contract" /
"model contract", §8d):

* camera matrices: ``GSCamera`` (reference ``sugar/sugar_scene/cameras.py:212-221``),
  ``getWorld2View2`` / ``getProjectionMatrix`` / ``focal2fov``
  (``sugar/gaussian_splatting/utils/graphics_utils.py:39-78``),
  ``load_cameras`` (``scene_representation.py:120-156``);
* trajectory JSON: ``dataset_utils/sample_custom_traj.py:44-108``;
* ``.ply`` vertex layout: ``sugar/gaussian_splatting/scene/gaussian_model.py:187-266``.

Everything here is numpy / torch on the CPU; nothing touches the GPU or the CUDA library.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


# ----------------------------------------------------------------------------- cameras
def focal2fov(focal: float, pixels: float) -> float:
    """graphics_utils.py:77-78."""
    return 2 * math.atan(pixels / (2 * focal))


def fov2focal(fov: float, pixels: float) -> float:
    """graphics_utils.py:74-75."""
    return pixels / (2 * math.tan(fov / 2))


def world2view(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """``getWorld2View2`` with translate=0, scale=1 (graphics_utils.py:39-50); returns fp32 [4,4]."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> torch.Tensor:
    """``getProjectionMatrix`` (graphics_utils.py:52-72): symmetric frustum, z_sign = +1."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class Camera:
    """The fields of ``GSCamera`` the rasterizer path reads (cameras.py:141-235)."""

    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # [4,4] fp32, = W2C transposed
    full_proj_transform: torch.Tensor  # [4,4] fp32, = view @ proj
    camera_center: torch.Tensor  # [3]
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)

    def to(self, device) -> "Camera":
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device), self.znear, self.zfar)

    def packed(self) -> torch.Tensor:
        """[37] fp32: view(16) | proj(16) | campos(3) | tanfovx | tanfovy — the per-camera payload the
        frame loop scatters."""
        return torch.cat([self.world_view_transform.reshape(-1).float().cpu(),
                          self.full_proj_transform.reshape(-1).float().cpu(),
                          self.camera_center.reshape(-1).float().cpu(),
                          torch.tensor([self.tanfovx, self.tanfovy], dtype=torch.float32)])


def camera_from_c2w(c2w: np.ndarray, fx: float, fy: float, w: int, h: int) -> Camera:
    """scene_representation.py:144-156 followed by the GSCamera constructor (cameras.py:212-221)."""
    w2c = np.linalg.inv(c2w)
    R = np.transpose(w2c[:3, :3])
    T = w2c[:3, 3]
    FovY = focal2fov(fy, h)
    FovX = focal2fov(fx, w)
    view = torch.tensor(world2view(R, T)).transpose(0, 1).contiguous()
    proj = projection_matrix(0.01, 100.0, FovX, FovY).transpose(0, 1).contiguous()
    full = (view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = view.inverse()[3, :3].contiguous()
    return Camera(w, h, FovX, FovY, view, full, center)


def _normalize(v, eps=1e-8):
    return v / (np.linalg.norm(v) + eps)


def rotm_from_lookat(lookat, up):
    """sample_custom_traj.py:44-50 (OpenCV camera-to-world rotation)."""
    z_axis = _normalize(lookat)
    x_axis = _normalize(np.cross(z_axis, up))
    y_axis = _normalize(np.cross(z_axis, x_axis))
    return np.array((x_axis, y_axis, z_axis)).T


def grid_half_sphere(radius=1.5, num_views=30, theta=None, phi_range=(0, 360)):
    """sample_custom_traj.py:53-66."""
    if theta is None:
        theta = np.deg2rad(np.array((0, 15, 30, 45, 60)))
    else:
        theta = np.deg2rad([theta])
    phi = np.deg2rad(np.linspace(phi_range[0], phi_range[1], num_views // len(theta) + 1)[:-1])
    theta, phi = np.meshgrid(theta, phi)
    theta = theta.flatten()
    phi = phi.flatten()
    x = np.cos(theta) * np.cos(phi) * radius
    y = np.cos(theta) * np.sin(phi) * radius
    z = np.sin(theta) * radius
    return np.stack((x, y, z), axis=-1)


def trajectory_dict(radius=4.0, num_views=300, theta=30.0, center=(0.0, 0.0, 0.0), w=1920, h=1080,
                    fov_x_deg=60.0, name="synthetic") -> Dict:
    """The ``custom_camera_path/*.json`` schema (sample_custom_traj.py:69-108)."""
    center = np.asarray(center, dtype=np.float64)
    fl = w / (2 * math.tan(math.radians(fov_x_deg) / 2))
    poses = []
    for t in grid_half_sphere(radius, num_views, theta, (0, 360)) + center:
        R = rotm_from_lookat(center - t, np.array([0, 0, 1]))
        c2w = np.vstack((np.hstack((R, t.reshape(3, 1))), np.array([0, 0, 0, 1])))
        poses.append(c2w)
    frames = [{"filename": "{:05d}.png".format(i), "transform_matrix": c2w.tolist()} for i, c2w in enumerate(poses)]
    return {"trajectory_name": name, "camera_model": "OPENCV", "fl_x": fl, "fl_y": fl, "cx": w / 2, "cy": h / 2,
            "w": w, "h": h, "frames": frames}


def cameras_from_trajectory(traj: Dict, downscale: float = 1.0) -> List[Camera]:
    """scene_representation.py:120-156: frames sorted by filename, cx/cy ignored by the 3DGS path."""
    fx, fy, w, h = traj["fl_x"], traj["fl_y"], traj["w"], traj["h"]
    if downscale > 1.0:
        h = round(h / downscale)
        w = round(w / downscale)
        fx = fx / downscale
        fy = fy / downscale
    c2w = dict(sorted((f["filename"], np.array(f["transform_matrix"])) for f in traj["frames"]))
    return [camera_from_c2w(m, fx, fy, w, h) for m in c2w.values()]


def save_trajectory(path: str, traj: Dict) -> None:
    with open(path, "w") as f:
        json.dump(traj, f, indent=4)


def load_trajectory(path: str) -> Dict:
    with open(path) as f:
        return json.load(f)


def lookat_camera(eye: Sequence[float], target: Sequence[float], w: int, h: int, fov_x_deg: float = 60.0,
                  up=(0.0, 0.0, 1.0), fov_y_deg: Optional[float] = None) -> Camera:
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    R = rotm_from_lookat(target - eye, np.asarray(up, dtype=np.float64))
    c2w = np.vstack((np.hstack((R, eye.reshape(3, 1))), np.array([0, 0, 0, 1])))
    fx = w / (2 * math.tan(math.radians(fov_x_deg) / 2))
    fy = fx if fov_y_deg is None else h / (2 * math.tan(math.radians(fov_y_deg) / 2))
    return camera_from_c2w(c2w, fx, fy, w, h)


# ----------------------------------------------------------------------------- synthetic Gaussians
def synthetic_gaussians(P: int, seed: int = 0, extent=(1.0, 1.0, 1.0), log_scale_mean: float = math.log(0.03),
                        log_scale_std: float = 0.4, opacity_mean: float = 1.0, opacity_std: float = 1.5,
                        sh_degree: int = 3, M: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Deterministic random scene with the distributions of SURVEY §8d.  Returns ACTIVATED tensors, i.e.
    what ``GaussianModel.get_*`` would hand to the rasterizer (gaussian_model.py:95-115): scales > 0,
    unit quaternions, opacities in (0,1) shaped [P,1], shs [P,M,3] coefficient-major."""
    g = torch.Generator().manual_seed(seed)
    M = (sh_degree + 1) ** 2 if M is None else M
    ext = torch.tensor(extent, dtype=torch.float32)
    means = (torch.rand(P, 3, generator=g) * 2 - 1) * ext
    scales = torch.exp(torch.randn(P, 3, generator=g) * log_scale_std + log_scale_mean)
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(torch.randn(P, 1, generator=g) * opacity_std + opacity_mean)
    shs = torch.randn(P, M, 3, generator=g) * 0.1
    shs[:, 0, :] = torch.randn(P, 3, generator=g) * 0.5
    return {"means3D": means.contiguous(), "scales": scales.contiguous(), "rotations": rotations.contiguous(),
            "opacities": opacities.contiguous(), "shs": shs.contiguous()}


def config1_scene():
    """BASELINE config 1: 10k Gaussians, one 256x256 camera."""
    g = synthetic_gaussians(10_000, seed=0, extent=(1, 1, 1), log_scale_mean=math.log(0.03), log_scale_std=0.4,
                            opacity_mean=1.0, opacity_std=1.5)
    cam = lookat_camera((0.0, -3.5, 0.0), (0, 0, 0), 256, 256, 60.0, fov_y_deg=60.0)
    return g, cam


def config3_scene(P: int = 3_000_000, seed: int = 1234):
    """BASELINE configs 3/4: 3M Gaussians, SH degree 3, over [-4,4]^2 x [-1,1]."""
    return synthetic_gaussians(P, seed=seed, extent=(4, 4, 1), log_scale_mean=math.log(0.006), log_scale_std=0.5,
                               opacity_mean=0.0, opacity_std=2.0)


def config2_raw(P: int = 1_000_000, seed: int = 1) -> Dict[str, torch.Tensor]:
    """BASELINE config 2 stand-in (the Garden .ply is not available offline): P = 1M, seed 1, the config-3
    distributions scaled to a 4-unit scene, as RAW (pre-activation) GaussianModel parameters — xyz [P,3], f_dc [P,1,3],
    f_rest [P,15,3], opacity [P,1] (logit), scaling [P,3] (log), rotation [P,4] (unnormalised) — i.e. what a 3DGS .ply stores."""
    g = torch.Generator().manual_seed(seed)
    ext = torch.tensor((2.0, 2.0, 0.5))
    return {"xyz": ((torch.rand(P, 3, generator=g) * 2 - 1) * ext).contiguous(),
            "scaling": (torch.randn(P, 3, generator=g) * 0.5 + math.log(0.005)).contiguous(),
            "rotation": torch.randn(P, 4, generator=g).contiguous(),
            "opacity": (torch.randn(P, 1, generator=g) * 2.0).contiguous(),
            "f_dc": (torch.randn(P, 1, 3, generator=g) * 0.5).contiguous(),
            "f_rest": (torch.randn(P, 15, 3, generator=g) * 0.1).contiguous()}


def config2_camera() -> "Camera":
    """One 1920x1080 camera, 60 degrees horizontal FoV, looking at the centre of the config-2 scene from its rim."""
    return lookat_camera((1.6, -2.6, 1.1), (0.0, 0.0, 0.0), 1920, 1080, 60.0)


# ----------------------------------------------------------------------------- .ply (3DGS layout)
def _ply_props(M: int) -> List[str]:
    props = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    props += ["f_rest_%d" % i for i in range(3 * (M - 1))]
    props += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    return props


def save_ply(path: str, xyz: np.ndarray, f_dc: np.ndarray, f_rest: np.ndarray, opacity_raw: np.ndarray,
             scale_raw: np.ndarray, rot_raw: np.ndarray) -> None:
    """Binary little-endian PLY in the layout of ``GaussianModel.save_ply`` (gaussian_model.py:187-223).
    f_dc [P,1,3], f_rest [P,M-1,3] are stored CHANNEL-major (transpose(1,2).flatten), values are the RAW
    (pre-activation) parameters."""
    P = xyz.shape[0]
    M = 1 + f_rest.shape[1]
    cols = [xyz, np.zeros_like(xyz), f_dc.transpose(0, 2, 1).reshape(P, -1), f_rest.transpose(0, 2, 1).reshape(P, -1),
            opacity_raw.reshape(P, 1), scale_raw, rot_raw]
    data = np.concatenate(cols, axis=1).astype("<f4")
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join("property float %s\n" % p for p in _ply_props(M)) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(data.tobytes())


def load_ply(path: str) -> Dict[str, np.ndarray]:
    """Single read of the vertex block (the reference loops per property through plyfile,
    gaussian_model.py:225-266).  Returns RAW parameters: xyz [P,3], f_dc [P,1,3], f_rest [P,M-1,3],
    opacity [P,1], scale [P,3], rot [P,4]."""
    with open(path, "rb") as f:
        names = []
        count = 0
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                count = int(line.split()[-1])
            elif line.startswith("property"):
                parts = line.split()
                if parts[1] not in ("float", "float32"):
                    raise ValueError("unsupported ply property type: " + line)
                names.append(parts[2])
            elif line == "end_header":
                break
        data = np.frombuffer(f.read(count * len(names) * 4), dtype="<f4").reshape(count, len(names))
    col = {n: i for i, n in enumerate(names)}
    xyz = data[:, [col["x"], col["y"], col["z"]]]
    f_dc = data[:, [col["f_dc_0"], col["f_dc_1"], col["f_dc_2"]]].reshape(count, 3, 1).transpose(0, 2, 1)
    rest_names = sorted((n for n in names if n.startswith("f_rest_")), key=lambda s: int(s.split("_")[-1]))
    f_rest = data[:, [col[n] for n in rest_names]].reshape(count, 3, len(rest_names) // 3).transpose(0, 2, 1)
    scale_names = sorted((n for n in names if n.startswith("scale_")), key=lambda s: int(s.split("_")[-1]))
    rot_names = sorted((n for n in names if n.startswith("rot")), key=lambda s: int(s.split("_")[-1]))
    return {"xyz": np.ascontiguousarray(xyz), "f_dc": np.ascontiguousarray(f_dc), "f_rest": np.ascontiguousarray(f_rest),
            "opacity": np.ascontiguousarray(data[:, [col["opacity"]]]),
            "scale": np.ascontiguousarray(data[:, [col[n] for n in scale_names]]),
            "rot": np.ascontiguousarray(data[:, [col[n] for n in rot_names]])}


def activate(raw: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    """The ``GaussianModel.get_*`` activations (gaussian_model.py:95-115)."""
    rot = torch.from_numpy(raw["rot"]).float()
    return {"means3D": torch.from_numpy(raw["xyz"]).float().contiguous(),
            "scales": torch.exp(torch.from_numpy(raw["scale"]).float()).contiguous(),
            "rotations": torch.nn.functional.normalize(rot).contiguous(),
            "opacities": torch.sigmoid(torch.from_numpy(raw["opacity"]).float()).contiguous(),
            "shs": torch.cat([torch.from_numpy(raw["f_dc"]).float(), torch.from_numpy(raw["f_rest"]).float()], dim=1).contiguous()}

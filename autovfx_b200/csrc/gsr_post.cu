// gsr_b200 — the per-Gaussian and per-pixel work the reference's render() wrappers do around the two rasterizer passes
// (sugar/gaussian_splatting/gaussian_renderer/__init__.py:83-218 = "GR/", utils/general_utils.py = "GU/"), as three
// HBM-streaming kernels instead of ~25 elementwise torch launches:
//
//   k_axis_normals   per-Gaussian shading normal = shortest axis of the Gaussian, flipped towards the camera
//                    (GaussianModel.get_normal, scene/gaussian_model.py:120-128; GU/:78-99,136-157), optionally remapped
//                    to [0,1] (GR/:147) — the colors_precomp of the reference's second pass
//   k_normal_maps    rendered normal image -> unit normals [H,W,3] (GR/:168-176) and the pseudo normal from the depth map
//                    (depth_pcd2normal + get_ray_directions, GR/:23-38,41-80,178-191)
//   k_pack_frame     8-bit hand-off of a finished frame: RGBA (torchvision.utils.save_image rounding), normal map and
//                    depth colormap index (scene_representation.py:424-438, sugar/render.py:18-22)
//
// Each torch op of the reference rounds once, so the arithmetic below uses explicit round-to-nearest intrinsics in the
// reference's operation order (no FMA contraction); reductions over 3-4 elements are summed left to right.
#include "gsr_common.cuh"

namespace gsr {

__device__ __forceinline__ float norm3_rn(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__global__ void __launch_bounds__(256) k_axis_normals(int P, const float* __restrict__ means3D, const float* __restrict__ scales,
                                                      const float* __restrict__ rotations, const float* __restrict__ campos, int remap01,
                                                      float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float s0 = scales[3 * (size_t)idx], s1 = scales[3 * (size_t)idx + 1], s2 = scales[3 * (size_t)idx + 2];
    // argsort(scales)[0]: index of the smallest scale (GU/:137); ties resolve to the lowest index
    int k = 0;
    float sm = s0;
    if (s1 < sm) { sm = s1; k = 1; }
    if (s2 < sm) { sm = s2; k = 2; }
    // build_rotation (GU/:78-99): normalise, then the k-th COLUMN of R
    float q0 = rotations[4 * (size_t)idx], q1 = rotations[4 * (size_t)idx + 1], q2 = rotations[4 * (size_t)idx + 2], q3 = rotations[4 * (size_t)idx + 3];
    const float qn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2)), __fmul_rn(q3, q3)));
    const float r = __fdiv_rn(q0, qn), x = __fdiv_rn(q1, qn), y = __fdiv_rn(q2, qn), z = __fdiv_rn(q3, qn);
    float n0, n1, n2;
    if (k == 0) {
        n0 = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
        n1 = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, y), __fmul_rn(r, z)));
        n2 = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(x, z), __fmul_rn(r, y)));
    } else if (k == 1) {
        n0 = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(x, y), __fmul_rn(r, z)));
        n1 = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z))));
        n2 = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(y, z), __fmul_rn(r, x)));
    } else {
        n0 = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, z), __fmul_rn(r, y)));
        n1 = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(y, z), __fmul_rn(r, x)));
        n2 = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));
    }
    // dir_pp_normalized (GR/:131-132) and flip_align_view (GU/:151-157): keep the axis if it faces the camera
    const float dx = __fsub_rn(means3D[3 * (size_t)idx], campos[0]), dy = __fsub_rn(means3D[3 * (size_t)idx + 1], campos[1]),
                dz = __fsub_rn(means3D[3 * (size_t)idx + 2], campos[2]);
    const float dn = norm3_rn(dx, dy, dz);
    const float vx = __fdiv_rn(dx, dn), vy = __fdiv_rn(dy, dn), vz = __fdiv_rn(dz, dn);
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(n0, -vx), __fmul_rn(n1, -vy)), __fmul_rn(n2, -vz));
    if (!(dot >= 0.0f)) { n0 = -n0; n1 = -n1; n2 = -n2; }
    const float nn = norm3_rn(n0, n1, n2);
    n0 = __fdiv_rn(n0, nn); n1 = __fdiv_rn(n1, nn); n2 = __fdiv_rn(n2, nn);
    if (remap01) {  // normal * 0.5 + 0.5 (GR/:147)
        n0 = __fadd_rn(__fmul_rn(n0, 0.5f), 0.5f); n1 = __fadd_rn(__fmul_rn(n1, 0.5f), 0.5f); n2 = __fadd_rn(__fmul_rn(n2, 0.5f), 0.5f);
    }
    out[3 * (size_t)idx] = n0; out[3 * (size_t)idx + 1] = n1; out[3 * (size_t)idx + 2] = n2;
}

// world-space point of pixel (x, y) at the rendered depth (GR/:41-80,185-190): directions @ c2w[:3,:3].T * depth + c2w[:3,3]
struct PseudoCam {
    float m[12];  // c2w rows 0..2 (row-major 3x4) of the matrix the reference calls c2w = world_view_transform.inverse()
    float fx, fy, cx, cy;
};
__device__ __forceinline__ float3 unproject(const PseudoCam& c, int x, int y, float depth) {
    const float d0 = __fdiv_rn(__fadd_rn(__fsub_rn((float)x, c.cx), 0.5f), c.fx);
    const float d1 = __fdiv_rn(__fadd_rn(__fsub_rn((float)y, c.cy), 0.5f), c.fy);
    float3 p;
    p.x = __fadd_rn(c.m[3], __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, c.m[0]), __fmul_rn(d1, c.m[1])), c.m[2]), depth));
    p.y = __fadd_rn(c.m[7], __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, c.m[4]), __fmul_rn(d1, c.m[5])), c.m[6]), depth));
    p.z = __fadd_rn(c.m[11], __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, c.m[8]), __fmul_rn(d1, c.m[9])), c.m[10]), depth));
    return p;
}

__global__ void __launch_bounds__(256) k_normal_maps(int W, int H, const float* __restrict__ normal_img /*[3,H,W] or null*/,
                                                     const float* __restrict__ depth /*[H,W] or null*/, const float* __restrict__ c2w,
                                                     float fx, float fy, float cx, float cy, float* __restrict__ out_normal /*[H,W,3]*/,
                                                     float* __restrict__ out_pseudo /*[H,W,3]*/) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t pid = (size_t)y * W + x, HW = (size_t)W * H;
    if (normal_img && out_normal) {  // (img - 0.5) * 2, then F.normalize(p=2, dim=-1, eps=1e-12)  (GR/:168-176)
        const float a = __fmul_rn(__fsub_rn(normal_img[pid], 0.5f), 2.0f), b = __fmul_rn(__fsub_rn(normal_img[HW + pid], 0.5f), 2.0f),
                    c = __fmul_rn(__fsub_rn(normal_img[2 * HW + pid], 0.5f), 2.0f);
        const float n = fmaxf(norm3_rn(a, b, c), 1e-12f);
        out_normal[3 * pid] = __fdiv_rn(a, n); out_normal[3 * pid + 1] = __fdiv_rn(b, n); out_normal[3 * pid + 2] = __fdiv_rn(c, n);
    }
    if (depth && out_pseudo) {  // depth_pcd2normal (GR/:23-38): central differences, zero border
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        if (x > 0 && y > 0 && x < W - 1 && y < H - 1) {
            PseudoCam c;
#pragma unroll
            for (int i = 0; i < 12; i++) c.m[i] = c2w[i];
            c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
            const float3 pr = unproject(c, x + 1, y, depth[pid + 1]), pl = unproject(c, x - 1, y, depth[pid - 1]);
            const float3 pt = unproject(c, x, y - 1, depth[pid - W]), pb = unproject(c, x, y + 1, depth[pid + W]);
            const float ax = __fsub_rn(pr.x, pl.x), ay = __fsub_rn(pr.y, pl.y), az = __fsub_rn(pr.z, pl.z);  // left_to_right
            const float bx = __fsub_rn(pt.x, pb.x), by = __fsub_rn(pt.y, pb.y), bz = __fsub_rn(pt.z, pb.z);  // bottom_to_top
            const float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by)), ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz)),
                        nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
            const float n = fmaxf(norm3_rn(nx, ny, nz), 1e-12f);
            o0 = __fdiv_rn(nx, n); o1 = __fdiv_rn(ny, n); o2 = __fdiv_rn(nz, n);
        }
        out_pseudo[3 * pid] = o0; out_pseudo[3 * pid + 1] = o1; out_pseudo[3 * pid + 2] = o2;
    }
}

// 8-bit hand-off.  rgba8: torchvision.utils.save_image = mul(255).add(0.5).clamp(0,255).to(uint8) on cat(rgb, alpha)
// (scene_representation.py:424-425, GR/:143); normal8: ((n + 1) / 2 * 255).astype(uint8) (scene_representation.py:433-436);
// depth8: (clip(depth / scale, 0, 1) * 255).astype(uint8), the index into the TURBO colormap (sugar/render.py:18-22).
__device__ __forceinline__ uint32_t q_save_image(float v) {
    const float t = fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f), 0.0f), 255.0f);
    return (uint32_t)t;  // NaN -> 0 like the clamp + cast on the GPU path
}
__global__ void __launch_bounds__(256) k_pack_frame(int W, int H, const float* __restrict__ rgb, const float* __restrict__ alpha,
                                                    const float* __restrict__ depth, const float* __restrict__ normal_hwc, float depth_scale,
                                                    uint32_t* __restrict__ rgba8, uint8_t* __restrict__ normal8, uint8_t* __restrict__ depth8) {
    const size_t HW = (size_t)W * H;
    const size_t pid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= HW) return;
    if (rgba8) {
        const uint32_t a = alpha ? q_save_image(alpha[pid]) : 255u;
        rgba8[pid] = q_save_image(rgb[pid]) | (q_save_image(rgb[HW + pid]) << 8) | (q_save_image(rgb[2 * HW + pid]) << 16) | (a << 24);
    }
    if (normal8) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = __fmul_rn(__fdiv_rn(__fadd_rn(normal_hwc[3 * pid + c], 1.0f), 2.0f), 255.0f);
            normal8[3 * pid + c] = (uint8_t)(int)v;  // numpy astype(uint8) truncates; the value is in [0, 255]
        }
    }
    if (depth8) {
        const float v = __fmul_rn(fminf(fmaxf(__fdiv_rn(depth[pid], depth_scale), 0.0f), 1.0f), 255.0f);
        depth8[pid] = (uint8_t)(int)v;
    }
}

int axis_normals_impl(int P, const float* means3D, const float* scales, const float* rotations, const float* campos, int remap01,
                      float* out, cudaStream_t st) {
    if (P < 0 || (P > 0 && (!means3D || !scales || !rotations || !campos || !out))) { set_error("gsr_axis_normals: bad arguments"); return GSR_ERR_INVALID; }
    if (P == 0) return GSR_OK;
    k_axis_normals<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, scales, rotations, campos, remap01, out);
    return check_launch("gsr_axis_normals", false, st);
}

int normal_maps_impl(int W, int H, const float* normal_img, const float* depth, const float* c2w, float fx, float fy, float cx, float cy,
                     float* out_normal, float* out_pseudo, cudaStream_t st) {
    if (W <= 0 || H <= 0) { set_error("gsr_normal_maps: bad size"); return GSR_ERR_INVALID; }
    if ((normal_img == nullptr) != (out_normal == nullptr) || (depth == nullptr) != (out_pseudo == nullptr) || (depth && !c2w)) {
        set_error("gsr_normal_maps: inputs and outputs must come in pairs");
        return GSR_ERR_INVALID;
    }
    k_normal_maps<<<dim3((W + 31) / 32, (H + 7) / 8), 256, 0, st>>>(W, H, normal_img, depth, c2w, fx, fy, cx, cy, out_normal, out_pseudo);
    return check_launch("gsr_normal_maps", false, st);
}

int pack_frame_impl(int W, int H, const float* rgb, const float* alpha, const float* depth, const float* normal_hwc, float depth_scale,
                    uint8_t* rgba8, uint8_t* normal8, uint8_t* depth8, cudaStream_t st) {
    if (W <= 0 || H <= 0) { set_error("gsr_pack_frame: bad size"); return GSR_ERR_INVALID; }
    if ((rgba8 && !rgb) || (normal8 && !normal_hwc) || (depth8 && (!depth || !(depth_scale > 0.0f)))) { set_error("gsr_pack_frame: missing input for a requested output"); return GSR_ERR_INVALID; }
    if (rgba8 && ((uintptr_t)rgba8 & 3)) { set_error("gsr_pack_frame: rgba8 must be 4-byte aligned"); return GSR_ERR_INVALID; }
    const size_t HW = (size_t)W * H;
    k_pack_frame<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(W, H, rgb, alpha, depth, normal_hwc, depth_scale, (uint32_t*)rgba8, normal8, depth8);
    return check_launch("gsr_pack_frame", false, st);
}

}  // namespace gsr

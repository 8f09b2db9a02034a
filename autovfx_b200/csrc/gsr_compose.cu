// gsr_b200 — Gaussian parameter activation and the per-frame rigid edit of inserted objects, in one streaming pass.
//
// The reference keeps RAW parameters (log-scales, unnormalised quaternions, logit opacities, SH split into dc / rest) and
// activates them with separate torch ops on every render call (scene/gaussian_model.py:95-115 = "GM/": exp, normalize,
// sigmoid, cat).  For edited scenes its frame loop additionally, per frame and per inserted object, re-reads the object
// .ply, applies transform_gaussians (gaussians_utils.py:88-125 = "GUt/": scale about a pivot, rotate, translate, compose
// quaternions, shift log-scales), concatenates everything into a new model (GUt/:71-84) and deep-copies the whole scene
// (scene_representation.py:357-371).  Here the scene is activated once into resident arrays with spare capacity and each
// object's raw parameters stay resident; one launch of k_compose per (frame, object) writes the transformed + activated
// object straight into the tail of the scene arrays the rasterizer reads.
//
// Arithmetic: one IEEE rounding per reference torch op, in the reference's order (explicit _rn intrinsics, no contraction);
// the [N,3]x[3,3] product and the 4-element norm are summed left to right.
#include "gsr_common.cuh"

namespace gsr {

struct ComposeParams {
    int N, M;                    // Gaussians, SH coefficients per channel in the OUTPUT rows (1 + rest coefficients)
    const float* xyz;            // [N,3]
    const float* f_dc;           // [N,1,3]
    const float* f_rest;         // [N,M-1,3] (ignored when M == 1)
    const float* opacity_raw;    // [N]
    const float* scaling_raw;    // [N,3]
    const float* rotation_raw;   // [N,4]
    float *means3D, *shs, *opacities, *scales, *rotations;
    int has_xform;
    gsr_object_xform x;
};

__global__ void __launch_bounds__(256) k_compose(const ComposeParams p) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < p.N) {
        float a0 = p.xyz[3 * (size_t)n], a1 = p.xyz[3 * (size_t)n + 1], a2 = p.xyz[3 * (size_t)n + 2];
        float s0 = p.scaling_raw[3 * (size_t)n], s1 = p.scaling_raw[3 * (size_t)n + 1], s2 = p.scaling_raw[3 * (size_t)n + 2];
        float q0 = p.rotation_raw[4 * (size_t)n], q1 = p.rotation_raw[4 * (size_t)n + 1], q2 = p.rotation_raw[4 * (size_t)n + 2],
              q3 = p.rotation_raw[4 * (size_t)n + 3];
        if (p.has_xform) {
            const gsr_object_xform& x = p.x;
            const float c0 = x.initial_center[0], c1 = x.initial_center[1], c2 = x.initial_center[2];
            // scale about the pivot (GUt/:99-103): (xyz - c) * s + c, then log-scale += log(s)
            a0 = __fadd_rn(__fmul_rn(__fsub_rn(a0, c0), x.scaling), c0);
            a1 = __fadd_rn(__fmul_rn(__fsub_rn(a1, c1), x.scaling), c1);
            a2 = __fadd_rn(__fmul_rn(__fsub_rn(a2, c2), x.scaling), c2);
            s0 = __fadd_rn(s0, x.log_scaling); s1 = __fadd_rn(s1, x.log_scaling); s2 = __fadd_rn(s2, x.log_scaling);
            // rotate about the pivot (GUt/:105-109): (xyz - c) @ R^T + c
            a0 = __fsub_rn(a0, c0); a1 = __fsub_rn(a1, c1); a2 = __fsub_rn(a2, c2);
            const float* R = x.rotation;
            const float b0 = __fmaf_rn(a2, R[2], __fmaf_rn(a1, R[1], __fmul_rn(a0, R[0])));
            const float b1 = __fmaf_rn(a2, R[5], __fmaf_rn(a1, R[4], __fmul_rn(a0, R[3])));
            const float b2 = __fmaf_rn(a2, R[8], __fmaf_rn(a1, R[7], __fmul_rn(a0, R[6])));
            // translate (GUt/:112-114): + (center - c)
            a0 = __fadd_rn(__fadd_rn(b0, c0), __fsub_rn(x.center[0], c0));
            a1 = __fadd_rn(__fadd_rn(b1, c1), __fsub_rn(x.center[1], c1));
            a2 = __fadd_rn(__fadd_rn(b2, c2), __fsub_rn(x.center[2], c2));
            // quaternion_multiply(matrix_to_quaternion(R), q) + standardize_quaternion (rotation_utils.py:113-150)
            const float aw = x.quat[0], ax = x.quat[1], ay = x.quat[2], az = x.quat[3];
            const float ow = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(aw, q0), __fmul_rn(ax, q1)), __fmul_rn(ay, q2)), __fmul_rn(az, q3));
            const float ox = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, q1), __fmul_rn(ax, q0)), __fmul_rn(ay, q3)), __fmul_rn(az, q2));
            const float oy = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(aw, q2), __fmul_rn(ax, q3)), __fmul_rn(ay, q0)), __fmul_rn(az, q1));
            const float oz = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(aw, q3), __fmul_rn(ax, q2)), __fmul_rn(ay, q1)), __fmul_rn(az, q0));
            const bool neg = ow < 0.0f;
            q0 = neg ? -ow : ow; q1 = neg ? -ox : ox; q2 = neg ? -oy : oy; q3 = neg ? -oz : oz;
        }
        p.means3D[3 * (size_t)n] = a0; p.means3D[3 * (size_t)n + 1] = a1; p.means3D[3 * (size_t)n + 2] = a2;
        // activations (GM/:95-115): exp, F.normalize (x / max(||x||, 1e-12)), sigmoid = 1 / (1 + exp(-x))
        p.scales[3 * (size_t)n] = expf(s0); p.scales[3 * (size_t)n + 1] = expf(s1); p.scales[3 * (size_t)n + 2] = expf(s2);
        const float qn = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2)), __fmul_rn(q3, q3))), 1e-12f);
        p.rotations[4 * (size_t)n] = __fdiv_rn(q0, qn); p.rotations[4 * (size_t)n + 1] = __fdiv_rn(q1, qn);
        p.rotations[4 * (size_t)n + 2] = __fdiv_rn(q2, qn); p.rotations[4 * (size_t)n + 3] = __fdiv_rn(q3, qn);
        p.opacities[n] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-p.opacity_raw[n])));
    }
    // SH rows: shs[n] = cat(f_dc[n], f_rest[n]) (GM/:107-110).  The block's 256 rows are copied as one flat, coalesced range.
    const size_t row = (size_t)3 * p.M, rest_row = row - 3;
    const size_t first = (size_t)blockIdx.x * blockDim.x;
    const size_t rows_here = min((size_t)blockDim.x, (size_t)p.N - first);
    const size_t total = rows_here * row;
    float* dst = p.shs + first * row;
    for (size_t j = threadIdx.x; j < total; j += blockDim.x) {
        const size_t r = j / row, k = j - r * row;
        dst[j] = k < 3 ? p.f_dc[(first + r) * 3 + k] : p.f_rest[(first + r) * rest_row + (k - 3)];
    }
}

int compose_impl(int N, int M, const float* xyz, const float* f_dc, const float* f_rest, const float* opacity_raw, const float* scaling_raw,
                 const float* rotation_raw, const gsr_object_xform* xform, float* means3D, float* shs, float* opacities, float* scales,
                 float* rotations, cudaStream_t st) {
    if (N < 0 || M < 1) { set_error("gsr_activate_gaussians: bad sizes N=%d M=%d", N, M); return GSR_ERR_INVALID; }
    if (N == 0) return GSR_OK;
    if (!xyz || !f_dc || (M > 1 && !f_rest) || !opacity_raw || !scaling_raw || !rotation_raw || !means3D || !shs || !opacities || !scales || !rotations) {
        set_error("gsr_activate_gaussians: null pointer");
        return GSR_ERR_INVALID;
    }
    ComposeParams p{};
    p.N = N; p.M = M; p.xyz = xyz; p.f_dc = f_dc; p.f_rest = f_rest; p.opacity_raw = opacity_raw; p.scaling_raw = scaling_raw;
    p.rotation_raw = rotation_raw; p.means3D = means3D; p.shs = shs; p.opacities = opacities; p.scales = scales; p.rotations = rotations;
    p.has_xform = xform != nullptr;
    if (xform) p.x = *xform;
    k_compose<<<(N + 255) / 256, 256, 0, st>>>(p);
    return check_launch("gsr_activate_gaussians", false, st);
}

}  // namespace gsr

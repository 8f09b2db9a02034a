// gsr_b200 backward pass.
//
// Replaces CudaRasterizer::Rasterizer::backward (DGR/cuda_rasterizer/rasterizer_impl.cu:343-446):
//   k_blend_backward     <- BACKWARD::render        (backward.cu:415-599)
//   k_gaussian_backward  <- computeCov2DCUDA + BACKWARD::preprocessCUDA (backward.cu:144-274, :346-412,
//                           with the SH backward :20-139 and the scale/rotation backward :278-341)
//
// Differences in mechanism (the mathematics and the per-pixel recursion are the reference's):
//   * the reference issues 10 global atomicAdds per contributing (pixel, Gaussian) pair; here each warp
//     (an 8x4 pixel footprint) reduces the 10 partial gradients with shuffles and issues one global
//     reduction per (warp, Gaussian, component);
//   * splats that cannot touch a warp's footprint are culled exactly as in the forward blend;
//   * the two per-Gaussian backward kernels are fused; SH rows are staged through shared memory with
//     coalesced accesses in both directions, and the kernel writes every output row itself (zeros for
//     culled Gaussians) so only the 48 B/Gaussian of atomically accumulated gradients need a memset
//     (the reference zero-fills all 304 B/Gaussian, rasterize_points.cu:158-168).
// Gradient sums are accumulated in a different order than the reference's atomics (which are themselves
// unordered), so parity is to a tolerance, not bitwise.
#include "gsr_common.cuh"

namespace gsr {

constexpr int BWD_THREADS = 256;

__device__ __forceinline__ float4 lds128b(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128b(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Sum 10 per-lane values over the warp with a reduce-scatter butterfly: at every level each lane keeps the
// half of the values its group is responsible for and ships the other half, so 5+3+2+1+1 = 12 shuffles replace
// 10 full butterflies (50).  Afterwards the (even) lane with `valid` holds the warp total of value `vid`.
__device__ __forceinline__ float reduce10(const float (&v)[10], int lane, int& vid, bool& valid) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float w[5], x[3], y[2];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float send = b4 ? v[i] : v[i + 5], keep = b4 ? v[i + 5] : v[i];
        w[i] = keep + __shfl_xor_sync(GSR_FULL, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float hi = (i + 3 < 5) ? w[(i + 3 < 5) ? i + 3 : 0] : 0.f;
        const float send = b3 ? w[i] : hi, keep = b3 ? hi : w[i];
        x[i] = keep + __shfl_xor_sync(GSR_FULL, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float hi = (i + 2 < 3) ? x[(i + 2 < 3) ? i + 2 : 0] : 0.f;
        const float send = b2 ? x[i] : hi, keep = b2 ? hi : x[i];
        y[i] = keep + __shfl_xor_sync(GSR_FULL, send, 4);
    }
    const float send = b1 ? y[0] : y[1], keep = b1 ? y[1] : y[0];
    float z = keep + __shfl_xor_sync(GSR_FULL, send, 2);
    z += __shfl_xor_sync(GSR_FULL, z, 1);
    valid = !(b2 && (b3 || b1)) && !(lane & 1);
    vid = (b4 ? 5 : 0) + (b3 ? (b1 ? 4 : 3) : (b2 ? 2 : (b1 ? 1 : 0)));
    return z;
}

constexpr int BWD_QCAP = 48;

// One CTA per tile; batches of the tile's list are walked from the back.
__global__ void __launch_bounds__(BWD_THREADS) k_blend_backward(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const float4* __restrict__ records, int W, int H, int gx,
    const float* __restrict__ bg, const float* __restrict__ accum_alphas, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths, const float* __restrict__ dL_dpixel_alphas,
    float* __restrict__ dL_dmean2D /*[P,3]*/, float* __restrict__ dL_dconic /*[P,4]*/, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dcolors /*[P,3]*/, float* __restrict__ dL_ddepths) {
    __shared__ __align__(16) float4 sRec[BWD_THREADS * 3];                    // staged batch, 48 B per splat
    __shared__ __align__(16) float4 sQ[(BWD_THREADS / 32) * BWD_QCAP * 3];    // per-warp survivor queues (back to front)
    __shared__ uint32_t sId[BWD_THREADS];
    __shared__ uint32_t s_wl[BWD_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int X0 = blockIdx.x * GSR_TILE + (warp & 1) * 8, Y0 = blockIdx.y * GSR_TILE + (warp >> 1) * 4;
    const int pxi = X0 + (lane & 7), pyi = Y0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    // warp-uniform values go through a broadcast so the compiler keeps them instead of re-deriving them from the ids in the loops
    const float fcx = __shfl_sync(GSR_FULL, (float)X0 + FOOT_HX, 0), fcy = __shfl_sync(GSR_FULL, (float)Y0 + FOOT_HY, 0);
    const size_t pid = (size_t)W * pyi + pxi, HW = (size_t)H * W;
    const uint32_t rec_base = __shfl_sync(GSR_FULL, (uint32_t)__cvta_generic_to_shared(sRec), 0);
    const uint32_t q_base = __shfl_sync(GSR_FULL, (uint32_t)__cvta_generic_to_shared(sQ) + (uint32_t)warp * (BWD_QCAP * 48), 0);
    const unsigned gt_mask = lane == 31 ? 0u : (0xffffffffu << (lane + 1));

    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    if (n == 0) return;

    // After reduce10 the warp total of gradient component `vid` sits in one fixed lane (vid depends on the lane only), so
    // each of those ten lanes keeps the array and stride its component goes to and adds the total straight to global memory
    // with a fire-and-forget reduction (RED.ADD.F32): one per (warp, splat, component).  (Combining the 8 warps of a tile in
    // shared memory first costs a compare-and-swap loop per add — shared memory has no native float add — plus two extra
    // barriers and a flush pass per batch.)
    float* my_dst;
    uint32_t my_stride;
    {
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
        const int vid = (b4 ? 5 : 0) + (b3 ? (b1 ? 4 : 3) : (b2 ? 2 : (b1 ? 1 : 0)));  // as in reduce10
        float* const dst[10] = {dL_dcolors, dL_dcolors + 1, dL_dcolors + 2, dL_ddepths, dL_dmean2D, dL_dmean2D + 1,
                                dL_dconic, dL_dconic + 1, dL_dconic + 3, dL_dopacity};
        const uint32_t strd[10] = {3, 3, 3, 1, 3, 3, 4, 4, 4, 1};
        my_dst = dst[0];
        my_stride = strd[0];
#pragma unroll
        for (int k = 1; k < 10; k++)
            if (vid == k) { my_dst = dst[k]; my_stride = strd[k]; }
    }

    const float T_final = inside ? (1 - accum_alphas[pid]) : 0;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pid] : 0;
    float accum_rec0 = 0, accum_rec1 = 0, accum_rec2 = 0, accum_red = 0, accum_rea = 0;
    float dLp0 = 0, dLp1 = 0, dLp2 = 0, dLd = 0, dLa = 0;
    if (inside) {
        dLp0 = dL_dpixels[pid]; dLp1 = dL_dpixels[HW + pid]; dLp2 = dL_dpixels[2 * HW + pid];
        dLd = dL_dpixel_depths[pid];
        dLa = dL_dpixel_alphas[pid];
    }
    float last_alpha = 0, last_c0 = 0, last_c1 = 0, last_c2 = 0, last_depth = 0;
    const float ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;  // backward.cu:488-489
    const float bg_dot_dpixel = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;

    // entries with 1-based position > the warp's furthest contributor are skipped by every lane
    uint32_t warp_last = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(GSR_FULL, warp_last, o));
    if (lane == 0) s_wl[warp] = warp_last;
    __syncthreads();
    uint32_t tile_last = 0;
#pragma unroll
    for (int k = 0; k < BWD_THREADS / 32; k++) tile_last = max(tile_last, s_wl[k]);
    if (tile_last == 0) return;  // nothing contributed anywhere in this tile

    int qn = 0;
    // replay the queued splats (back to front) for this lane's pixel: backward.cu:494-597
    auto drain = [&](int batch) {
        __syncwarp();
        uint32_t qa = q_base;
        for (int k = 0; k < qn; k++, qa += 48) {
            const float4 A = lds128b(qa), B = lds128b(qa + 16), Cc = lds128b(qa + 32);
            const uint32_t pos = __float_as_uint(Cc.w);  // 1-based position in the tile list
            float g[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // dcolor[3], ddepth, dmean2D.x, .y, dconic.x, .y, .w, dopacity
            bool contrib = false;
            if (pos <= last_contributor) {
                const float2 d = {A.x - pixx, A.y - pixy};
                const float power = -0.5f * (A.z * d.x * d.x + B.x * d.y * d.y) - A.w * d.x * d.y;
                if (!(power > 0.0f)) {
                    const float G = exp(power);
                    const float alpha = min(0.99f, B.y * G);
                    if (!(alpha < 1.0f / 255.0f)) {
                        contrib = true;
                        T = T / (1.f - alpha);
                        const float dchannel_dcolor = alpha * T;
                        float dL_dalpha = 0.0f;
                        accum_rec0 = last_alpha * last_c0 + (1.f - last_alpha) * accum_rec0;
                        last_c0 = Cc.x;
                        dL_dalpha += (Cc.x - accum_rec0) * dLp0;
                        g[0] = dchannel_dcolor * dLp0;
                        accum_rec1 = last_alpha * last_c1 + (1.f - last_alpha) * accum_rec1;
                        last_c1 = Cc.y;
                        dL_dalpha += (Cc.y - accum_rec1) * dLp1;
                        g[1] = dchannel_dcolor * dLp1;
                        accum_rec2 = last_alpha * last_c2 + (1.f - last_alpha) * accum_rec2;
                        last_c2 = Cc.z;
                        dL_dalpha += (Cc.z - accum_rec2) * dLp2;
                        g[2] = dchannel_dcolor * dLp2;
                        const float dep = B.z;
                        accum_red = last_alpha * last_depth + (1.f - last_alpha) * accum_red;
                        last_depth = dep;
                        dL_dalpha += (dep - accum_red) * dLd;
                        g[3] = dchannel_dcolor * dLd;
                        accum_rea = last_alpha + (1.f - last_alpha) * accum_rea;
                        dL_dalpha += (1 - accum_rea) * dLa;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                        const float dL_dG = B.y * dL_dalpha;
                        const float gdx = G * d.x, gdy = G * d.y;
                        const float dG_ddelx = -gdx * A.z - gdy * A.w;
                        const float dG_ddely = -gdy * B.x - gdx * A.w;
                        g[4] = dL_dG * dG_ddelx * ddelx_dx;
                        g[5] = dL_dG * dG_ddely * ddely_dy;
                        g[6] = -0.5f * gdx * d.x * dL_dG;
                        g[7] = -0.5f * gdx * d.y * dL_dG;
                        g[8] = -0.5f * gdy * d.y * dL_dG;
                        g[9] = G * dL_dalpha;
                    }
                }
            }
            if (__any_sync(GSR_FULL, contrib)) {
                int vid;
                bool valid;
                const float z = reduce10(g, lane, vid, valid);
                (void)vid;
                if (valid && z != 0.0f) atomicAdd(my_dst + (size_t)sId[(int)pos - 1 - batch * BWD_THREADS] * my_stride, z);
            }
        }
        qn = 0;
        __syncwarp();
    };

    for (int b = (int)((tile_last - 1) / BWD_THREADS); b >= 0; b--) {
        const int cnt = min(BWD_THREADS, n - b * BWD_THREADS);
        __syncthreads();  // every warp is done with the previous batch's records
        if (tid < cnt) {
            const uint32_t id = point_list[range.x + b * BWD_THREADS + tid];
            const float4* r = records + 3 * (size_t)id;
            float4 rc = r[2];
            rc.w = __uint_as_float((uint32_t)(b * BWD_THREADS + tid + 1));
            const uint32_t sa = rec_base + (uint32_t)tid * 48;
            sts128b(sa, r[0]); sts128b(sa + 16, r[1]); sts128b(sa + 32, rc);
            sId[tid] = id;
        }
        __syncthreads();

        if ((uint32_t)(b * BWD_THREADS) < warp_last) {
            for (int base = ((cnt - 1) / 32) * 32; base >= 0; base -= 32) {
                const int s = base + lane;
                const uint32_t sa = rec_base + (uint32_t)s * 48;
                bool keep = false;
                float4 A, B;
                if (s < cnt && (uint32_t)(b * BWD_THREADS + s + 1) <= warp_last) {
                    A = lds128b(sa); B = lds128b(sa + 16);
                    keep = footprint_may_touch(A.x - fcx, A.y - fcy, A.z, A.w, B.x, B.w);
                }
                const unsigned mask = __ballot_sync(GSR_FULL, keep);
                if (mask) {
                    if (keep) {  // later list positions (higher lanes) are replayed first
                        const uint32_t qa = q_base + (uint32_t)(qn + __popc(mask & gt_mask)) * 48;
                        sts128b(qa, A); sts128b(qa + 16, B); sts128b(qa + 32, lds128b(sa + 32));
                    }
                    qn += __popc(mask);
                    if (qn > BWD_QCAP - 32) drain(b);
                }
            }
            if (qn) drain(b);
        }
    }
}

// =====================================================================================================
// per-Gaussian backward (fused computeCov2DCUDA + preprocessCUDA backward)
// =====================================================================================================
struct GBParams {
    int P, D, M, W, H;
    float scale_modifier, tanfovx, tanfovy, h_x, h_y;
    const float *means3D, *shs, *scales, *rotations, *cov3Ds, *view, *proj, *campos;
    const int* radii;
    const uint8_t* clamped;
    const float *dL_dmean2D, *dL_dconic, *dL_dcolor, *dL_ddepth;
    float *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscale, *dL_drot;
};

constexpr int GB_THREADS = 128;
constexpr int GB_STRIDE = 49;  // 48 SH floats per row, odd stride -> conflict-free

__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {  // auxiliary.h:106-118
    float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
    float3 o;
    o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return o;
}

// SH backward for one Gaussian (backward.cu:20-139): reads its coefficients from `sh` and overwrites the
// same row with dL/dsh (entries beyond the active degree become 0).  Returns dL/dmean from the view direction.
__device__ float3 sh_backward(int deg, float* sh, float3 pos, const float* campos, unsigned clamp_bits, float3 dL_dcolor) {
    float3 dir_orig = {pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]};
    float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    float dRGB[3] = {dL_dcolor.x * ((clamp_bits & 1u) ? 0.f : 1.f), dL_dcolor.y * ((clamp_bits & 2u) ? 0.f : 1.f),
                     dL_dcolor.z * ((clamp_bits & 4u) ? 0.f : 1.f)};
    float ddir[3] = {0, 0, 0};
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    float w[16];
    w[0] = SH_C0;
    w[1] = -SH_C1 * y; w[2] = SH_C1 * z; w[3] = -SH_C1 * x;
    w[4] = SH_C2_0 * xy; w[5] = SH_C2_1 * yz; w[6] = SH_C2_2 * (2.f * zz - xx - yy); w[7] = SH_C2_3 * xz; w[8] = SH_C2_4 * (xx - yy);
    w[9] = SH_C3_0 * y * (3.f * xx - yy); w[10] = SH_C3_1 * xy * z; w[11] = SH_C3_2 * y * (4.f * zz - xx - yy);
    w[12] = SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy); w[13] = SH_C3_4 * x * (4.f * zz - xx - yy);
    w[14] = SH_C3_5 * z * (xx - yy); w[15] = SH_C3_6 * x * (xx - 3.f * yy);
    const int ncoef = (deg + 1) * (deg + 1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
#define S(k) sh[(k)*3 + c]
        float dx = 0, dy = 0, dz = 0;
        if (deg > 0) {
            dx = -SH_C1 * S(3);
            dy = -SH_C1 * S(1);
            dz = SH_C1 * S(2);
            if (deg > 1) {
                dx += SH_C2_0 * y * S(4) + SH_C2_2 * 2.f * -x * S(6) + SH_C2_3 * z * S(7) + SH_C2_4 * 2.f * x * S(8);
                dy += SH_C2_0 * x * S(4) + SH_C2_1 * z * S(5) + SH_C2_2 * 2.f * -y * S(6) + SH_C2_4 * 2.f * -y * S(8);
                dz += SH_C2_1 * y * S(5) + SH_C2_2 * 2.f * 2.f * z * S(6) + SH_C2_3 * x * S(7);
                if (deg > 2) {
                    dx += (SH_C3_0 * S(9) * 3.f * 2.f * xy + SH_C3_1 * S(10) * yz + SH_C3_2 * S(11) * -2.f * xy + SH_C3_3 * S(12) * -3.f * 2.f * xz +
                           SH_C3_4 * S(13) * (-3.f * xx + 4.f * zz - yy) + SH_C3_5 * S(14) * 2.f * xz + SH_C3_6 * S(15) * 3.f * (xx - yy));
                    dy += (SH_C3_0 * S(9) * 3.f * (xx - yy) + SH_C3_1 * S(10) * xz + SH_C3_2 * S(11) * (-3.f * yy + 4.f * zz - xx) +
                           SH_C3_3 * S(12) * -3.f * 2.f * yz + SH_C3_4 * S(13) * -2.f * xy + SH_C3_5 * S(14) * -2.f * yz + SH_C3_6 * S(15) * -3.f * 2.f * xy);
                    dz += (SH_C3_1 * S(10) * xy + SH_C3_2 * S(11) * 4.f * 2.f * yz + SH_C3_3 * S(12) * 3.f * (2.f * zz - xx - yy) +
                           SH_C3_4 * S(13) * 4.f * 2.f * xz + SH_C3_5 * S(14) * (xx - yy));
                }
            }
        }
        ddir[0] += dx * dRGB[c];
        ddir[1] += dy * dRGB[c];
        ddir[2] += dz * dRGB[c];
#pragma unroll
        for (int k = 0; k < 16; k++) S(k) = k < ncoef ? w[k] * dRGB[c] : 0.f;
#undef S
    }
    return dnormvdv3(dir_orig, make_float3(ddir[0], ddir[1], ddir[2]));
}

// M16: shs has exactly 16 coefficients (48 floats, 16-byte aligned rows): rows move as float4 with compile-time indexing
template <bool M16>
__global__ void __launch_bounds__(GB_THREADS) k_gaussian_backward(const GBParams p) {
    __shared__ CamConsts cam;
    __shared__ float stage[GB_THREADS * GB_STRIDE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) cam.view[tid] = p.view[tid];
    else if (tid < 32) cam.proj[tid - 16] = p.proj[tid - 16];
    else if (tid < 35) cam.campos[tid - 32] = p.campos[tid - 32];
    __syncthreads();
    const int idx = blockIdx.x * GB_THREADS + tid;
    const bool valid = idx < p.P;
    const bool vis = valid && p.radii[idx] > 0;
    const float* view = cam.view;
    const float* proj = cam.proj;

    // ---- stage the SH rows of visible Gaussians (first 16 coefficients) ----
    float* wstage = stage + warp * 32 * GB_STRIDE;
    const size_t gbase = (size_t)(blockIdx.x * GB_THREADS + warp * 32);
    const size_t row_floats = (size_t)p.M * 3;
    const int nf = p.shs ? min(48, (int)row_floats) : 0;
    const unsigned vismask = __ballot_sync(GSR_FULL, vis);
    if (p.shs) {
        if (M16) {  // 12 float4 per row, consecutive lanes fetch consecutive 16-byte parts
#pragma unroll
            for (int it = 0; it < 12; it++) {
                const int item = it * 32 + lane;
                const int gl = item / 12, part = item - gl * 12;
                if ((vismask >> gl) & 1u) {
                    const float4 v = reinterpret_cast<const float4*>(p.shs + (gbase + gl) * 48)[part];
                    float* d = wstage + gl * GB_STRIDE + part * 4;
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        } else {
            for (int it = 0; it < nf; it++) {
                const int item = it * 32 + lane;
                const int gl = item / nf, part = item - gl * nf;
                if ((vismask >> gl) & 1u) wstage[gl * GB_STRIDE + part] = p.shs[(gbase + gl) * row_floats + part];
            }
        }
        __syncwarp();
    }

    float3 dmean = {0, 0, 0};
    float dcov[6] = {0, 0, 0, 0, 0, 0};
    float3 dscale = {0, 0, 0};
    float4 drot = {0, 0, 0, 0};
    if (vis) {
        const float3 mean = {p.means3D[3 * (size_t)idx], p.means3D[3 * (size_t)idx + 1], p.means3D[3 * (size_t)idx + 2]};
        // ---------------- computeCov2DCUDA (backward.cu:144-274) ----------------
        {
            float c3[6];
#pragma unroll
            for (int k = 0; k < 6; k++) c3[k] = p.cov3Ds[6 * (size_t)idx + k];
            const float3 dL_dconic = {p.dL_dconic[4 * (size_t)idx], p.dL_dconic[4 * (size_t)idx + 1], p.dL_dconic[4 * (size_t)idx + 3]};
            float3 t = xform4x3(mean, view);
            const float limx = 1.3f * p.tanfovx, limy = 1.3f * p.tanfovy;
            const float txtz = t.x / t.z, tytz = t.y / t.z;
            t.x = min(limx, max(-limx, txtz)) * t.z;
            t.y = min(limy, max(-limy, tytz)) * t.z;
            const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
            const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;
            m3 J = m3_make(p.h_x / t.z, 0.0f, -(p.h_x * t.x) / (t.z * t.z), 0.0f, p.h_y / t.z, -(p.h_y * t.y) / (t.z * t.z), 0, 0, 0);
            m3 Wm = m3_make(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
            m3 Vrk = m3_make(c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]);
            m3 T = m3_mul(Wm, J);
            m3 cov2D = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
            const float a = cov2D.m[0][0] + 0.3f, b = cov2D.m[0][1], c = cov2D.m[1][1] + 0.3f;
            const float denom = a * c - b * b;
            float dL_da = 0, dL_db = 0, dL_dc = 0;
            const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
#define TT(c_, r_) T.m[c_][r_]
#define VV(c_, r_) Vrk.m[c_][r_]
            if (denom2inv != 0) {
                dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
                dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
                dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);
                dcov[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
                dcov[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
                dcov[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
                dcov[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 1) * dL_dc;
                dcov[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 2) * dL_dc;
                dcov[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db + 2 * TT(1, 1) * TT(1, 2) * dL_dc;
            }
            const float dL_dT00 = 2 * (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_da + (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_db;
            const float dL_dT01 = 2 * (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_da + (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_db;
            const float dL_dT02 = 2 * (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_da + (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_db;
            const float dL_dT10 = 2 * (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_dc + (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_db;
            const float dL_dT11 = 2 * (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_dc + (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_db;
            const float dL_dT12 = 2 * (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_dc + (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_db;
#undef TT
#undef VV
            const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
            const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
            const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
            const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;
            const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
            const float dL_dtx = x_grad_mul * -p.h_x * tz2 * dL_dJ02;
            const float dL_dty = y_grad_mul * -p.h_y * tz2 * dL_dJ12;
            const float dL_dtz = -p.h_x * tz2 * dL_dJ00 - p.h_y * tz2 * dL_dJ11 + (2 * p.h_x * t.x) * tz3 * dL_dJ02 + (2 * p.h_y * t.y) * tz3 * dL_dJ12;
            // transformVec4x3Transpose (auxiliary.h:89-97)
            dmean.x = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
            dmean.y = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
            dmean.z = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;
        }
        // ---------------- preprocessCUDA backward (backward.cu:346-412) ----------------
        {
            const float3 m = mean;
            const float4 m_hom = xform4x4(m, proj);
            const float m_w = 1.0f / (m_hom.w + 0.0000001f);
            const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
            const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
            const float d2x = p.dL_dmean2D[3 * (size_t)idx], d2y = p.dL_dmean2D[3 * (size_t)idx + 1];
            float3 dL_dmean;
            dL_dmean.x = (proj[0] * m_w - proj[3] * mul1) * d2x + (proj[1] * m_w - proj[3] * mul2) * d2y;
            dL_dmean.y = (proj[4] * m_w - proj[7] * mul1) * d2x + (proj[5] * m_w - proj[7] * mul2) * d2y;
            dL_dmean.z = (proj[8] * m_w - proj[11] * mul1) * d2x + (proj[9] * m_w - proj[11] * mul2) * d2y;
            dmean.x += dL_dmean.x; dmean.y += dL_dmean.y; dmean.z += dL_dmean.z;
            const float mul3 = view[2] * m.x + view[6] * m.y + view[10] * m.z + view[14];
            const float dd = p.dL_ddepth[idx];
            dmean.x += (view[2] - view[3] * mul3) * dd;
            dmean.y += (view[6] - view[7] * mul3) * dd;
            dmean.z += (view[10] - view[11] * mul3) * dd;
            if (p.shs) {
                const float3 dcol = {p.dL_dcolor[3 * (size_t)idx], p.dL_dcolor[3 * (size_t)idx + 1], p.dL_dcolor[3 * (size_t)idx + 2]};
                const float3 dm = sh_backward(p.D, wstage + lane * GB_STRIDE, m, cam.campos, p.clamped[idx], dcol);
                dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
            }
            if (p.scales) {  // computeCov3D backward (backward.cu:278-341)
                const float sx = p.scales[3 * (size_t)idx], sy = p.scales[3 * (size_t)idx + 1], sz = p.scales[3 * (size_t)idx + 2];
                const float r = p.rotations[4 * (size_t)idx], x = p.rotations[4 * (size_t)idx + 1], y = p.rotations[4 * (size_t)idx + 2], z = p.rotations[4 * (size_t)idx + 3];
                m3 R = m3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                               2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                               2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
                const float s0 = p.scale_modifier * sx, s1 = p.scale_modifier * sy, s2 = p.scale_modifier * sz;
                m3 S = m3_make(s0, 0.0f, 0.0f, 0.0f, s1, 0.0f, 0.0f, 0.0f, s2);
                m3 Mm = m3_mul(S, R);
                m3 dL_dSigma = m3_make(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                       0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
                m3 M2;
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int rr = 0; rr < 3; rr++) M2.m[c][rr] = 2.0f * Mm.m[c][rr];
                m3 dL_dM = m3_mul(M2, dL_dSigma);
                m3 Rt = m3_t(R), dMt = m3_t(dL_dM);
                dscale.x = Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2];
                dscale.y = Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2];
                dscale.z = Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2];
#pragma unroll
                for (int rr = 0; rr < 3; rr++) { dMt.m[0][rr] *= s0; dMt.m[1][rr] *= s1; dMt.m[2][rr] *= s2; }
#define Dm(c_, r_) dMt.m[c_][r_]
                drot.x = 2 * z * (Dm(0, 1) - Dm(1, 0)) + 2 * y * (Dm(2, 0) - Dm(0, 2)) + 2 * x * (Dm(1, 2) - Dm(2, 1));
                drot.y = 2 * y * (Dm(1, 0) + Dm(0, 1)) + 2 * z * (Dm(2, 0) + Dm(0, 2)) + 2 * r * (Dm(1, 2) - Dm(2, 1)) - 4 * x * (Dm(2, 2) + Dm(1, 1));
                drot.z = 2 * x * (Dm(1, 0) + Dm(0, 1)) + 2 * r * (Dm(2, 0) - Dm(0, 2)) + 2 * z * (Dm(1, 2) + Dm(2, 1)) - 4 * y * (Dm(2, 2) + Dm(0, 0));
                drot.w = 2 * r * (Dm(0, 1) - Dm(1, 0)) + 2 * x * (Dm(2, 0) + Dm(0, 2)) + 2 * y * (Dm(1, 2) + Dm(2, 1)) - 4 * z * (Dm(1, 1) + Dm(0, 0));
#undef Dm
            }
        }
    }
    if (valid) {
        p.dL_dmeans3D[3 * (size_t)idx] = dmean.x; p.dL_dmeans3D[3 * (size_t)idx + 1] = dmean.y; p.dL_dmeans3D[3 * (size_t)idx + 2] = dmean.z;
#pragma unroll
        for (int k = 0; k < 6; k++) p.dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
        if (p.dL_dscale) { p.dL_dscale[3 * (size_t)idx] = dscale.x; p.dL_dscale[3 * (size_t)idx + 1] = dscale.y; p.dL_dscale[3 * (size_t)idx + 2] = dscale.z; }
        if (p.dL_drot) { p.dL_drot[4 * (size_t)idx] = drot.x; p.dL_drot[4 * (size_t)idx + 1] = drot.y; p.dL_drot[4 * (size_t)idx + 2] = drot.z; p.dL_drot[4 * (size_t)idx + 3] = drot.w; }
    }
    // ---- write dL/dsh rows: coalesced, zeros for culled Gaussians and for coefficients beyond 16 ----
    if (p.shs && p.dL_dsh) {
        __syncwarp();
        const int valid_rows = min(32, p.P - (int)gbase);
        if (M16) {
#pragma unroll
            for (int it = 0; it < 12; it++) {
                const int item = it * 32 + lane;
                const int gl = item / 12, part = item - gl * 12;
                if (gl < valid_rows) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((vismask >> gl) & 1u) {
                        const float* s = wstage + gl * GB_STRIDE + part * 4;
                        v = make_float4(s[0], s[1], s[2], s[3]);
                    }
                    reinterpret_cast<float4*>(p.dL_dsh + (gbase + gl) * 48)[part] = v;
                }
            }
        } else {
            const int rf = (int)row_floats;
            for (int item = lane; item < valid_rows * rf; item += 32) {
                const int gl = item / rf, part = item - gl * rf;
                float v = 0.f;
                if (((vismask >> gl) & 1u) && part < nf) v = wstage[gl * GB_STRIDE + part];
                p.dL_dsh[gbase * row_floats + item] = v;
            }
        }
    }
}

int backward_impl(const gsr_frame* f, const gsr_workspace* ws, const int32_t* radii, const float* out_alpha, const float* dL_dc,
                  const float* dL_dd, const float* dL_da, const gsr_grads* g, cudaStream_t st) {
    if (!f || !ws || !g) { set_error("gsr_backward: null argument"); return GSR_ERR_INVALID; }
    const bool debug = f->debug != 0;
    const size_t P = (size_t)f->P;
    if (!g->dL_dmeans2D || !g->dL_dconic || !g->dL_dopacity || !g->dL_dcolors || !g->dL_ddepths || !g->dL_dmeans3D || !g->dL_dcov3D) {
        set_error("gsr_backward: null gradient buffer");
        return GSR_ERR_INVALID;
    }
    if (f->shs && !g->dL_dsh) { set_error("gsr_backward: dL_dsh missing"); return GSR_ERR_INVALID; }
    if (f->scales && (!g->dL_dscales || !g->dL_drotations)) { set_error("gsr_backward: dL_dscales/dL_drotations missing"); return GSR_ERR_INVALID; }
    // accumulated gradients start from zero
    cudaMemsetAsync(g->dL_dmeans2D, 0, 12 * P, st);
    cudaMemsetAsync(g->dL_dconic, 0, 16 * P, st);
    cudaMemsetAsync(g->dL_dopacity, 0, 4 * P, st);
    cudaMemsetAsync(g->dL_dcolors, 0, 12 * P, st);
    cudaMemsetAsync(g->dL_ddepths, 0, 4 * P, st);
    if (P == 0) return check_launch("gsr_backward(P=0)", debug, st);
    if (!radii || !out_alpha || !dL_dc || !dL_dd || !dL_da) { set_error("gsr_backward: null input"); return GSR_ERR_INVALID; }
    const ImageLayout il(f->W, f->H);
    const GeomLayout gl(P);
    if (ws->image_bytes < il.total || ws->geom_bytes < gl.total || !ws->binning) { set_error("gsr_backward: workspace too small"); return GSR_ERR_WORKSPACE; }
    const BinLayout bl(BinLayout::capacity_of(ws->binning_bytes));
    char* img = (char*)ws->image; char* geo = (char*)ws->geom; char* bin = (char*)ws->binning;
    const int D = f->D < 0 ? 0 : (f->D > 3 ? 3 : f->D);

    k_blend_backward<<<dim3(il.gx, il.gy), BWD_THREADS, 0, st>>>(
        (const uint2*)(img + il.ranges), (const uint32_t*)(bin + bl.point_list), (const float4*)(geo + gl.records), f->W, f->H, il.gx, f->bg,
        out_alpha, (const uint32_t*)(img + il.n_contrib), dL_dc, dL_dd, dL_da, g->dL_dmeans2D, g->dL_dconic, g->dL_dopacity, g->dL_dcolors,
        g->dL_ddepths);
    int rc = check_launch("gsr_backward/blend", debug, st);
    if (rc) return rc;

    GBParams gp;
    gp.P = f->P; gp.D = D; gp.M = f->M; gp.W = f->W; gp.H = f->H;
    gp.scale_modifier = f->scale_modifier; gp.tanfovx = f->tanfovx; gp.tanfovy = f->tanfovy;
    gp.h_y = f->H / (2.0f * f->tanfovy); gp.h_x = f->W / (2.0f * f->tanfovx);
    gp.means3D = f->means3D; gp.shs = f->shs; gp.scales = f->scales; gp.rotations = f->rotations;
    gp.cov3Ds = f->cov3D_precomp ? f->cov3D_precomp : (const float*)(geo + gl.cov3D);
    gp.view = f->viewmatrix; gp.proj = f->projmatrix; gp.campos = f->campos;
    gp.radii = radii; gp.clamped = (const uint8_t*)(geo + gl.clamped);
    gp.dL_dmean2D = g->dL_dmeans2D; gp.dL_dconic = g->dL_dconic; gp.dL_dcolor = g->dL_dcolors; gp.dL_ddepth = g->dL_ddepths;
    gp.dL_dmeans3D = g->dL_dmeans3D; gp.dL_dcov3D = g->dL_dcov3D; gp.dL_dsh = g->dL_dsh;
    gp.dL_dscale = f->scales ? g->dL_dscales : nullptr; gp.dL_drot = f->scales ? g->dL_drotations : nullptr;
    const bool m16 = f->shs && f->M == 16 && (((uintptr_t)f->shs | (uintptr_t)g->dL_dsh) & 15) == 0;
    if (m16) k_gaussian_backward<true><<<(f->P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, st>>>(gp);
    else k_gaussian_backward<false><<<(f->P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, st>>>(gp);
    // gradient buffers the reference leaves at zero for the absent parametrisation
    if (!f->scales) {
        if (g->dL_dscales) cudaMemsetAsync(g->dL_dscales, 0, 12 * P, st);
        if (g->dL_drotations) cudaMemsetAsync(g->dL_drotations, 0, 16 * P, st);
    }
    return check_launch("gsr_backward/gaussian", debug, st);
}

}  // namespace gsr

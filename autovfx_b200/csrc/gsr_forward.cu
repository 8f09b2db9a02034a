// gsr_b200 forward pass: preprocess -> tile histogram/scan -> instance emission -> per-tile depth sort -> blend.
//
// Replaces CudaRasterizer::Rasterizer::forward (DGR/cuda_rasterizer/rasterizer_impl.cu:197-339) and the
// kernels it drives (forward.cu:155-256 preprocessCUDA, rasterizer_impl.cu:70-138 duplicateWithKeys /
// identifyTileRanges, CUB InclusiveSum + DeviceRadixSort, forward.cu:261-378 renderCUDA).
//
// Pipeline differences (results are identical):
//   * one 48-byte record per visible Gaussian {x,y,conic.a,conic.b | conic.c,opacity,depth,tau | r,g,b,-}
//     replaces the reference's five SoA arrays, so the blend gathers 3 aligned float4 per instance;
//   * SH coefficients are staged into shared memory with coalesced 16-byte cp.async by each warp, only for
//     the Gaussians that survived culling, and read back conflict-free (row stride 13 float4);
//   * binning is a two-level sort: per-tile histogram (atomics, in preprocess) -> exclusive scan over the
//     tiles (gives the ranges and R on the device, no host round trip) -> scatter of (depth bits, id) pairs
//     into the tile's bucket -> one CTA per tile sorts its bucket by (depth bits, id) in shared memory.
//     The resulting order is exactly the reference's stable radix sort of (tile | depth) keys, because the
//     reference emits every (tile, Gaussian) pair once in ascending Gaussian id (rasterizer_impl.cu:98-108);
//   * the blend kernel culls splats per 8x4-pixel warp footprint (one splat per lane, ballot) before the
//     per-pixel evaluation, which itself uses the reference's fp32 expressions.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "gsr_common.cuh"
#include "gsr_packed.cuh"

namespace gsr {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
int check_launch(const char* what, bool debug, cudaStream_t stream) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GSR_ERR_CUDA;
    }
    return GSR_OK;
}

// =====================================================================================================
// Kernel 1: preprocess
// =====================================================================================================
struct PreParams {
    int P, D, M, W, H, gx, gy;
    float scale_modifier, tanfovx, tanfovy, focal_x, focal_y;
    int prefiltered, for_backward, rot_vec, tight;
    const float *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp, *view, *proj, *campos;
    float4* records;
    float* cov3D;
    uint8_t* clamped;
    int* radii;
    uint32_t* tile_count;
    uint32_t* tile_big;
    uint32_t* ranks;
    gsr_counters* counters;
};

constexpr int PRE_THREADS = 128;

__host__ __device__ constexpr int sh_nf(int deg) { return 3 * (deg + 1) * (deg + 1); }
__host__ __device__ constexpr int sh_nv(int deg, bool win) { return (sh_nf(deg) + 3) / 4 + (win ? 1 : 0); }  // float4 per staged row
__host__ __device__ constexpr int sh_stride(int deg, bool vec, bool win = false) {
    // vec: rows of nv float4, padded so that (stride/4) is odd -> conflict-free LDS.128 across 8 lanes
    // scalar: odd number of words -> conflict-free LDS.32
    return vec ? (sh_nv(deg, win) % 2 == 0 ? (sh_nv(deg, win) + 1) * 4 : sh_nv(deg, win) * 4) : (sh_nf(deg) | 1);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// TMA bulk copy (cp.async.bulk, SASS UBLKCP) of one contiguous row global -> shared, completion on an mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// SH -> RGB for one Gaussian from its staged row (forward.cu:20-71).  Like cov3d_ref_rounding below, the
// roundings are pinned to the instruction sequence nvcc 12.9 emits for the reference (SASS of oracle/_ref):
// every coefficient C*poly(dir) is built from separate multiplies (only 3xx-yy, 4zz-xx, 2zz-3xx-3yy, xx-3yy use
// an FMA), and each term is accumulated with one FMA, in the reference's order.
template <int DEG>
__device__ __forceinline__ void sh_eval(const float* sh, float3 pos, const float* campos, float* rgb, unsigned& clamp_bits) {
    const float dx = __fsub_rn(pos.x, campos[0]), dy = __fsub_rn(pos.y, campos[1]), dz = __fsub_rn(pos.z, campos[2]);
    const float len = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
    const float x = __fdiv_rn(dx, len), y = __fdiv_rn(dy, len), z = __fdiv_rn(dz, len);
    float k[16];
    k[0] = SH_C0;
    if (DEG > 0) {
        k[1] = -__fmul_rn(y, SH_C1);
        k[2] = __fmul_rn(z, SH_C1);
        k[3] = -__fmul_rn(x, SH_C1);
    }
    if (DEG > 1) {
        const float xy = __fmul_rn(y, x), yz = __fmul_rn(z, y), zz = __fmul_rn(z, z), xx = __fmul_rn(x, x), yy = __fmul_rn(y, y);
        const float xz = __fmul_rn(z, x), zz2 = __fadd_rn(zz, zz), xmy = __fsub_rn(xx, yy);
        k[4] = __fmul_rn(xy, SH_C2_0);
        k[5] = __fmul_rn(yz, SH_C2_1);
        k[6] = __fmul_rn(__fsub_rn(__fsub_rn(zz2, xx), yy), SH_C2_2);
        k[7] = __fmul_rn(xz, SH_C2_3);
        k[8] = __fmul_rn(xmy, SH_C2_4);
        if (DEG > 2) {
            const float f = __fsub_rn(__fmaf_rn(zz, 4.0f, -xx), yy);  // 4zz - xx - yy
            k[9] = __fmul_rn(__fmul_rn(y, SH_C3_0), __fmaf_rn(xx, 3.0f, -yy));
            k[10] = __fmul_rn(__fmul_rn(xy, SH_C3_1), z);
            k[11] = __fmul_rn(__fmul_rn(y, SH_C3_2), f);
            k[12] = __fmul_rn(__fmul_rn(z, SH_C3_3), __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2)));
            k[13] = __fmul_rn(__fmul_rn(x, SH_C3_4), f);
            k[14] = __fmul_rn(__fmul_rn(z, SH_C3_5), xmy);
            k[15] = __fmul_rn(__fmul_rn(x, SH_C3_6), __fmaf_rn(yy, -3.0f, xx));
        }
    }
    constexpr int NC = (DEG + 1) * (DEG + 1);
    clamp_bits = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float res = __fmul_rn(sh[c], k[0]);
#pragma unroll
        for (int i = 1; i < NC; i++) res = __fmaf_rn(k[i], sh[i * 3 + c], res);
        const float v = __fadd_rn(res, 0.5f);
        const bool neg = v < 0.f;  // == (res < -0.5f)
        if (neg) clamp_bits |= 1u << c;
        rgb[c] = neg ? 0.0f : v;
    }
}

// 3D covariance from scale / rotation (forward.cu:118-152: Sigma = (S R)^T (S R), quaternion not normalised).
// The roundings (which product of each sum is fused into an FMA, which terms multiply the structural zeros of
// S) are pinned with intrinsics to what nvcc 12.9 emits for the reference's GLM expression on sm_100a
// (read from the SASS of oracle/_ref), because the compiler's contraction choices depend on common
// sub-expression sharing and cannot be reproduced by writing "the same" C++ expression in another kernel.
// A 1-ulp difference here changes conic -> alpha -> the T < 1e-4 termination of a pixel now and then.
__device__ __forceinline__ void cov3d_ref_rounding(float sx, float sy, float sz, float mod, float4 q, float* c3) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x);
    const float A02 = __fmaf_rn(r, y, xz), A20 = __fmaf_rn(-r, y, xz);
    const float rz = __fmul_rn(r, z);
    const float A12 = __fmaf_rn(y, z, -rx), A21 = __fmaf_rn(y, z, rx);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float A01 = __fmaf_rn(x, y, -rz), A10 = __fmaf_rn(x, y, rz);
    const float B22 = __fmaf_rn(x, x, yy), B00 = __fadd_rn(yy, zz), B11 = __fmaf_rn(x, x, zz);
    const float s0 = __fmul_rn(mod, sx), s1 = __fmul_rn(mod, sy), s2 = __fmul_rn(mod, sz);
    // rotation matrix, R<c><r> = column c, row r
    const float R00 = __fsub_rn(1.f, __fadd_rn(B00, B00)), R01 = __fadd_rn(A01, A01), R02 = __fadd_rn(A02, A02);
    const float R10 = __fadd_rn(A10, A10), R11 = __fsub_rn(1.f, __fadd_rn(B11, B11)), R12 = __fadd_rn(A12, A12);
    const float R20 = __fadd_rn(A20, A20), R21 = __fadd_rn(A21, A21), R22 = __fsub_rn(1.f, __fadd_rn(B22, B22));
    // M = S * R with S = diag(s0, s1, s2): M<c><r> = S[0][r] R[c][0] + S[1][r] R[c][1] + S[2][r] R[c][2]
    const float z0 = __fmul_rn(0.f, R21), z00 = __fmul_rn(0.f, R00), z11 = __fmul_rn(0.f, R11);
    const float uM22 = __fmaf_rn(0.f, R20, z0), uM20 = __fmaf_rn(s0, R20, z0), uM21 = __fmaf_rn(0.f, R20, __fmul_rn(s1, R21));
    const float uM01 = __fmaf_rn(s1, R01, z00), uM02 = __fmaf_rn(0.f, R01, z00), uM00 = __fmaf_rn(0.f, R01, __fmul_rn(s0, R00));
    const float uM10 = __fmaf_rn(s0, R10, z11), uM11 = __fmaf_rn(0.f, R10, __fmul_rn(s1, R11)), uM12 = __fmaf_rn(0.f, R10, z11);
    const float M20 = __fmaf_rn(0.f, R22, uM20), M21 = __fmaf_rn(0.f, R22, uM21), M22 = __fmaf_rn(s2, R22, uM22);
    const float M00 = __fmaf_rn(0.f, R02, uM00), M01 = __fmaf_rn(0.f, R02, uM01), M02 = __fmaf_rn(s2, R02, uM02);
    const float M10 = __fmaf_rn(0.f, R12, uM10), M11 = __fmaf_rn(0.f, R12, uM11), M12 = __fmaf_rn(s2, R12, uM12);
    // Sigma = M^T M, upper triangle
    c3[0] = __fmaf_rn(M02, M02, __fmaf_rn(M00, M00, __fmul_rn(M01, M01)));
    c3[1] = __fmaf_rn(M02, M12, __fmaf_rn(M00, M10, __fmul_rn(M01, M11)));
    c3[2] = __fmaf_rn(M02, M22, __fmaf_rn(M00, M20, __fmul_rn(M01, M21)));
    c3[3] = __fmaf_rn(M12, M12, __fmaf_rn(M10, M10, __fmul_rn(M11, M11)));
    c3[4] = __fmaf_rn(M12, M22, __fmaf_rn(M10, M20, __fmul_rn(M11, M21)));
    c3[5] = __fmaf_rn(M22, M22, __fmaf_rn(M20, M20, __fmul_rn(M21, M21)));
}

// DEG = -1: colours are precomputed.  VEC: SH rows are 16-byte aligned (M % 4 == 0) -> 16-byte staging, either one TMA
// bulk copy per visible Gaussian issued by its own lane (BULK) or coalesced cp.async by the whole warp.  WIN (with VEC and
// BULK): rows are only 4-byte aligned (M = 25, the SuGaR storage: 300-byte rows) — each lane bulk-copies the 16-byte aligned
// window that contains its coefficients (one extra float4) and evaluates from its row's offset inside the window.
template <int DEG, bool VEC, bool BULK, bool WIN = false>
__global__ void __launch_bounds__(PRE_THREADS, 8) k_preprocess(const PreParams p) {
    constexpr int DG = DEG < 0 ? 0 : DEG;
    constexpr int NF = sh_nf(DG);
    constexpr int STRIDE = sh_stride(DG, VEC, WIN);
    __shared__ CamConsts cam;
    __shared__ __align__(16) float stage[DEG < 0 ? 4 : PRE_THREADS * STRIDE];
    __shared__ __align__(8) unsigned long long stage_bar[PRE_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (BULK && lane == 0) {
        mbar_init((uint32_t)__cvta_generic_to_shared(&stage_bar[warp]), 1u);
        mbar_fence_init();
    }
    if (tid < 16) cam.view[tid] = p.view[tid];
    else if (tid < 32) cam.proj[tid - 16] = p.proj[tid - 16];
    else if (tid < 35) cam.campos[tid - 32] = p.campos[tid - 32];
    __syncthreads();

    const int idx = blockIdx.x * PRE_THREADS + tid;
    const bool valid = idx < p.P;
    bool vis = false;
    int radius = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float3 mean = {0, 0, 0};
    float px = 0, py = 0, depth = 0, con_a = 0, con_b = 0, con_c = 0;
    float c3[6] = {0, 0, 0, 0, 0, 0};

    // all per-Gaussian inputs are requested up front (one memory round trip instead of three dependent ones);
    // the few near-culled Gaussians pay for 32 unused bytes
    float opacity = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    float4 q = make_float4(0, 0, 0, 0);
    if (valid) {
        mean = make_float3(p.means3D[3 * (size_t)idx], p.means3D[3 * (size_t)idx + 1], p.means3D[3 * (size_t)idx + 2]);
        opacity = p.opacities[idx];
        if (p.cov3D_precomp != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; k++) c3[k] = p.cov3D_precomp[6 * (size_t)idx + k];
        } else {
            sx = p.scales[3 * (size_t)idx]; sy = p.scales[3 * (size_t)idx + 1]; sz = p.scales[3 * (size_t)idx + 2];
            if (p.rot_vec) q = reinterpret_cast<const float4*>(p.rotations)[idx];
            else q = make_float4(p.rotations[4 * (size_t)idx], p.rotations[4 * (size_t)idx + 1], p.rotations[4 * (size_t)idx + 2], p.rotations[4 * (size_t)idx + 3]);
        }
        // near cull (auxiliary.h:139-164): only view-space z is tested
        float4 p_hom = xform4x4(mean, cam.proj);
        float p_w = 1.0f / (p_hom.w + 0.0000001f);
        float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
        float3 p_view = xform4x3(mean, cam.view);
        if (p_view.z <= 0.2f) {
            if (p.prefiltered) p.counters->trapped = 1;
        } else {
            // 3D covariance (forward.cu:118-152)
            if (p.cov3D_precomp == nullptr) cov3d_ref_rounding(sx, sy, sz, p.scale_modifier, q, c3);
            // EWA 2D covariance (forward.cu:74-113)
            float3 t = xform4x3(mean, cam.view);
            const float limx = 1.3f * p.tanfovx, limy = 1.3f * p.tanfovy;
            const float txtz = t.x / t.z, tytz = t.y / t.z;
            t.x = min(limx, max(-limx, txtz)) * t.z;
            t.y = min(limy, max(-limy, tytz)) * t.z;
            m3 J = m3_make(p.focal_x / t.z, 0.0f, -(p.focal_x * t.x) / (t.z * t.z),
                           0.0f, p.focal_y / t.z, -(p.focal_y * t.y) / (t.z * t.z),
                           0, 0, 0);
            m3 Wm = m3_make(cam.view[0], cam.view[4], cam.view[8], cam.view[1], cam.view[5], cam.view[9],
                            cam.view[2], cam.view[6], cam.view[10]);
            m3 T = m3_mul(Wm, J);
            m3 Vrk = m3_make(c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]);
            m3 cov = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
            cov.m[0][0] += 0.3f;
            cov.m[1][1] += 0.3f;
            const float cx = cov.m[0][0], cy = cov.m[0][1], cz = cov.m[1][1];
            // conic, radius, tile rectangle (forward.cu:217-237)
            float det = (cx * cz - cy * cy);
            if (det != 0.0f) {
                float det_inv = 1.f / det;
                con_a = cz * det_inv; con_b = -cy * det_inv; con_c = cx * det_inv;
                float mid = 0.5f * (cx + cz);
                float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
                float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
                float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
                px = ndc2pix(p_proj.x, p.W);
                py = ndc2pix(p_proj.y, p.H);
                tile_rect(px, py, (int)my_radius, p.gx, p.gy, x0, y0, x1, y1);
                if ((x1 - x0) * (y1 - y0) != 0) {
                    vis = true;
                    radius = (int)my_radius;
                    depth = p_view.z;
                }
            }
        }
    }
    if (!vis) { x0 = y0 = x1 = y1 = 0; }

    // Per-tile histogram.  Gaussians touching <= 8 tiles take a ranked ticket per tile (atomic with return; the
    // results are only needed at the end of the kernel, so the round trips overlap the SH work) and store the
    // ranks for k_emit, which then needs no atomics.  Larger rectangles are counted separately, warp-cooperatively.
    // With p.tight, a tile of the reference's rectangle only receives an instance if the splat can reach
    // alpha >= 1/255 at one of its pixel centres (tile_may_touch) — instances the reference creates but skips at
    // every pixel are never emitted (opt-in: the per-tile lists then differ from the reference's, the images do not).
    const float tau = footprint_tau(opacity);
    uint32_t rk[8];
    const int rect_w = x1 - x0, rect_n = rect_w * (y1 - y0);
    if (rect_n > 0 && rect_n <= 8) {
        int tx = x0, ty = y0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < rect_n) {
                rk[k] = 0xffffffffu;
                if (!p.tight || tile_may_touch(px, py, con_a, con_b, con_c, tau, tx, ty)) rk[k] = atomicAdd(&p.tile_count[ty * p.gx + tx], 1u);
                if (++tx == x1) { tx = x0; ty++; }
            }
        }
    }
    {
        const bool big = rect_n > 8;
        uint32_t* tb = p.tile_big;
        const int tight = p.tight;
        const uint32_t pay[6] = {__float_as_uint(px), __float_as_uint(py), __float_as_uint(con_a), __float_as_uint(con_b),
                                 __float_as_uint(con_c), __float_as_uint(tau)};
        for_each_tile<0, 6>(big ? x0 : 0, big ? y0 : 0, big ? x1 : 0, big ? y1 : 0, p.gx, pay, [&](int tile, int tx, int ty, const uint32_t(&o)[6]) {
            if (!tight || tile_may_touch(__uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]),
                                         __uint_as_float(o[4]), __uint_as_float(o[5]), tx, ty))
                atomicAdd(&tb[tile], 1u);
        });
    }

    // colour
    float rgb[3] = {0, 0, 0};
    unsigned clamp_bits = 0;
    if constexpr (DEG < 0) {
        if (vis) {
            rgb[0] = p.colors_precomp[3 * (size_t)idx];
            rgb[1] = p.colors_precomp[3 * (size_t)idx + 1];
            rgb[2] = p.colors_precomp[3 * (size_t)idx + 2];
        }
    } else {
        // Stage the SH rows of this warp's 32 Gaussians: flat work list (Gaussian, part), consecutive lanes
        // fetch consecutive 16-byte (or 4-byte) parts -> coalesced; rows of culled Gaussians are skipped.
        const unsigned vismask = __ballot_sync(GSR_FULL, vis);
        float* wstage = stage + warp * 32 * STRIDE;
        const size_t gbase = (size_t)(blockIdx.x * PRE_THREADS + warp * 32);
        const size_t row_floats = (size_t)p.M * 3;
        int win_off = 0;  // floats between the start of the staged window and the row's first coefficient (WIN only)
        if (VEC && BULK) {
            constexpr int NV = sh_nv(DG, WIN);
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar[warp]);
            if (vismask) {
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)__popc(vismask) * NV * 16u);
                if (vis) {
                    const float* src = p.shs + (gbase + lane) * row_floats;
                    if (WIN) {
                        win_off = (int)(((uintptr_t)src & 15u) >> 2);
                        src -= win_off;
                    }
                    bulk_g2s((uint32_t)__cvta_generic_to_shared(wstage + lane * STRIDE), src, NV * 16u, bar);
                }
                mbar_wait(bar, 0u);
            }
        } else if (VEC) {
            constexpr int NV = (NF + 3) / 4;
#pragma unroll
            for (int it = 0; it < NV; it++) {
                const int item = it * 32 + lane;
                const int gl = item / NV, part = item - gl * NV;
                if ((vismask >> gl) & 1u) cp_async16(wstage + gl * STRIDE + part * 4, p.shs + (gbase + gl) * row_floats + part * 4);
            }
            cp_async_wait_all();
        } else {
#pragma unroll 4
            for (int it = 0; it < NF; it++) {
                const int item = it * 32 + lane;
                const int gl = item / NF, part = item - gl * NF;
                if ((vismask >> gl) & 1u) wstage[gl * STRIDE + part] = p.shs[(gbase + gl) * row_floats + part];
            }
        }
        __syncwarp();
        if (vis) sh_eval<DG>(wstage + lane * STRIDE + win_off, mean, cam.campos, rgb, clamp_bits);
    }

    if (valid) {
        p.radii[idx] = radius;
        if (vis) {
            float4* rec = p.records + 3 * (size_t)idx;
            rec[0] = make_float4(px, py, con_a, con_b);
            rec[1] = make_float4(con_c, opacity, depth, tau);
            rec[2] = make_float4(rgb[0], rgb[1], rgb[2], log2f(opacity));  // .w: log2(opacity) for the default (fast-alpha) blend
            if (rect_n <= 8) {
                uint4* rr = reinterpret_cast<uint4*>(p.ranks + 8 * (size_t)idx);
                rr[0] = make_uint4(rk[0], rk[1], rk[2], rk[3]);
                if (rect_n > 4) rr[1] = make_uint4(rk[4], rk[5], rk[6], rk[7]);
            }
            if (p.for_backward) {
                if (p.cov3D_precomp == nullptr) {
#pragma unroll
                    for (int k = 0; k < 6; k++) p.cov3D[6 * (size_t)idx + k] = c3[k];
                }
                p.clamped[idx] = (uint8_t)clamp_bits;
            }
        }
    }
    const int nvis = __syncthreads_count(vis);
    if (tid == 0 && nvis) atomicAdd(&p.counters->num_visible, (uint32_t)nvis);
}

// =====================================================================================================
// Kernel 2: exclusive scan over the per-tile counts -> ranges, R, overflow flag (single CTA)
// =====================================================================================================
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ tile_big,
                                                    uint32_t* __restrict__ tile_fill, uint2* __restrict__ ranges,
                                                    gsr_counters* counters, int tiles, uint32_t capacity) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t warp_max[32];
    __shared__ uint32_t chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t carry = 0, lmax = 0;
    for (int base = 0; base < tiles; base += 1024) {  // chunks of 1024 consecutive tiles: coalesced loads and stores
        const int t = base + tid;
        const uint32_t cs = t < tiles ? tile_count[t] : 0u, c = cs + (t < tiles ? tile_big[t] : 0u);
        lmax = max(lmax, c);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(GSR_FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t s = warp_sum[lane];
            uint32_t si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(GSR_FULL, si, o);
                if (lane >= o) si += v;
            }
            warp_sum[lane] = si - s;  // exclusive
            if (lane == 31) chunk_total = si;
        }
        __syncthreads();
        const uint32_t start = carry + warp_sum[warp] + (incl - c);
        if (t < tiles) {
            ranges[t] = c ? make_uint2(start, start + c) : make_uint2(0u, 0u);
            tile_fill[t] = start + cs;  // absolute cursor of the tile's un-ranked (large-rectangle) instances
        }
        carry += chunk_total;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(GSR_FULL, lmax, o));
    if (lane == 0) warp_max[warp] = lmax;
    __syncthreads();
    if (warp == 0) {
        uint32_t m = warp_max[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(GSR_FULL, m, o));
        if (lane == 0) {
            counters->num_rendered = carry;
            counters->overflow = carry > capacity ? 1u : 0u;
            counters->max_tile = m;
        }
    }
}

// =====================================================================================================
// Kernel 3: scatter one (depth bits, Gaussian id) pair per (Gaussian, tile) instance into the tile's bucket
// (the work of duplicateWithKeys, rasterizer_impl.cu:70-111; the tile id is implicit in the bucket)
// =====================================================================================================
template <bool TIGHT>
__global__ void __launch_bounds__(256) k_emit(int P, int gx, int gy, const int* __restrict__ radii,
                                              const float4* __restrict__ records, const uint32_t* __restrict__ ranks,
                                              const uint2* __restrict__ ranges, uint32_t* __restrict__ tile_fill,
                                              uint2* __restrict__ pairs, const gsr_counters* __restrict__ counters) {
    if (counters->overflow) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    uint32_t dbits = 0;
    float4 r0 = make_float4(0, 0, 0, 0), r1 = r0;
    if (idx < P) {
        const int r = radii[idx];
        if (r > 0) {
            r0 = records[3 * (size_t)idx];
            if (TIGHT) {
                r1 = records[3 * (size_t)idx + 1];
                dbits = __float_as_uint(r1.z);
            } else {
                dbits = __float_as_uint(records[3 * (size_t)idx + 1].z);
            }
            tile_rect(r0.x, r0.y, r, gx, gy, x0, y0, x1, y1);
        }
    }
    // <= 8 tiles: the in-tile rank of every instance was drawn by k_preprocess -> plain scatter, no atomics
    // (rank 0xffffffff = tile culled by the tight-tile test)
    const int w = x1 - x0, cnt = w * (y1 - y0);
    if (cnt > 0 && cnt <= 8) {
        const uint4* rr = reinterpret_cast<const uint4*>(ranks + 8 * (size_t)idx);
        const uint4 ra = rr[0];
        uint4 rb = make_uint4(0, 0, 0, 0);
        if (cnt > 4) rb = rr[1];
        const uint32_t rk[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
        const uint2 pr = make_uint2((uint32_t)idx, dbits);  // little endian: u64 = (depth bits << 32) | id
        int tx = x0, ty = y0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < cnt) {
                if (!TIGHT || rk[k] != 0xffffffffu) pairs[ranges[ty * gx + tx].x + rk[k]] = pr;
                if (++tx == x1) { tx = x0; ty++; }
            }
        }
    }
    // > 8 tiles: walked by the whole warp, positions from the per-tile cursor initialised by k_tile_scan; the
    // tight-tile test is re-evaluated on the same stored values k_preprocess used (bitwise same decision)
    const bool big = cnt > 8;
    if (TIGHT) {
        const uint32_t pay[8] = {(uint32_t)idx, dbits, __float_as_uint(r0.x), __float_as_uint(r0.y), __float_as_uint(r0.z),
                                 __float_as_uint(r0.w), __float_as_uint(r1.x), __float_as_uint(r1.w)};
        for_each_tile<0, 8>(big ? x0 : 0, big ? y0 : 0, big ? x1 : 0, big ? y1 : 0, gx, pay, [&](int tile, int tx, int ty, const uint32_t(&o)[8]) {
            if (tile_may_touch(__uint_as_float(o[2]), __uint_as_float(o[3]), __uint_as_float(o[4]), __uint_as_float(o[5]),
                               __uint_as_float(o[6]), __uint_as_float(o[7]), tx, ty))
                pairs[atomicAdd(&tile_fill[tile], 1u)] = make_uint2(o[0], o[1]);
        });
    } else {
        const uint32_t pay[2] = {(uint32_t)idx, dbits};
        for_each_tile<0, 2>(big ? x0 : 0, big ? y0 : 0, big ? x1 : 0, big ? y1 : 0, gx, pay, [&](int tile, int, int, const uint32_t(&o)[2]) {
            pairs[atomicAdd(&tile_fill[tile], 1u)] = make_uint2(o[0], o[1]);
        });
    }
}

// =====================================================================================================
// Kernel 4: per-tile sort of the bucket by (depth bits, id) -> point_list
// Normalised bitonic network (all compare-exchanges ascending), valid for any n by skipping partners >= n.
// =====================================================================================================
constexpr int SORT_THREADS = 256;
constexpr int SORT_CAP = 4096;  // u64 entries in shared memory (32 KB)

template <typename T>
__device__ __forceinline__ void cmpx(T* a, uint32_t i, uint32_t l) {
    const unsigned long long x = a[i], y = a[l];
    if (x > y) { a[i] = y; a[l] = x; }
}
// one "flip" step of block size k over the first N slots (N power of two), entries >= n are virtual +inf
__device__ __forceinline__ void step_flip(unsigned long long* a, uint32_t n, uint32_t N, uint32_t k) {
    const uint32_t half = k >> 1;
    for (uint32_t t = threadIdx.x; t < (N >> 1); t += SORT_THREADS) {
        const uint32_t i = ((t & ~(half - 1)) << 1) | (t & (half - 1));
        const uint32_t l = i ^ (k - 1);
        if (l < n) cmpx(a, i, l);
    }
}
__device__ __forceinline__ void step_j(unsigned long long* a, uint32_t n, uint32_t N, uint32_t j) {
    for (uint32_t t = threadIdx.x; t < (N >> 1); t += SORT_THREADS) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t l = i + j;
        if (l < n) cmpx(a, i, l);
    }
}
__device__ __forceinline__ uint32_t next_pow2(uint32_t n) { return n <= 1 ? 1u : 1u << (32 - __clz(n - 1)); }

// full sort of a[0..n) (n <= SORT_CAP) in shared memory
__device__ void sort_smem(unsigned long long* s, uint32_t n) {
    const uint32_t N = next_pow2(n);
    for (uint32_t k = 2; k <= N; k <<= 1) {
        step_flip(s, n, N, k);
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            step_j(s, n, N, j);
            __syncthreads();
        }
    }
}

// ---- bucket sort for n <= SORT_CAP (the common case) -----------------------------------------------------
// The keys of one tile are (depth bits << 32 | id) with depths spread over [zmin, zmax] of the tile.  A monotone
// linear quantisation of the depth into B >= n/2 buckets (float subtract, multiply by a positive constant and
// truncation are all monotone) puts ~1 key in each bucket, so the sort is: histogram (shared atomics, the
// returned value is the key's slot in its bucket) -> exclusive scan -> scatter -> per-bucket insertion sort on
// the full 64-bit key, O(n) work instead of the O(n log^2 n) network.  The result is the unique ascending
// order, whatever order the atomics happened in.  Buckets holding more than SORT_BUCKET_MAX keys (all depths
// equal, or extreme clustering) trigger the generic network on the same shared array instead.
constexpr int SORT_BUCKETS = 2048;
constexpr int SORT_BUCKET_MAX = 24;

__device__ void sort_bucket(const unsigned long long* __restrict__ g, uint32_t n, uint32_t* __restrict__ out,
                            unsigned long long* __restrict__ gkeep, unsigned long long* s, uint32_t* hist /*[SORT_BUCKETS+1]*/) {
    __shared__ uint32_t red_min[SORT_THREADS / 32], red_max[SORT_THREADS / 32], wsum[SORT_THREADS / 32];
    __shared__ int fallback;
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t B = min((uint32_t)SORT_BUCKETS, max(32u, next_pow2(n)));
    // the keys are read three times from global memory (L2 hits after the first pass) instead of being held in
    // up to 32 registers per thread: min/max, bucket + ticket, scatter
    uint32_t dmin = 0xffffffffu, dmax = 0u;
    for (uint32_t i = t; i < n; i += SORT_THREADS) {
        const uint32_t d = (uint32_t)(g[i] >> 32);
        dmin = min(dmin, d);
        dmax = max(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = min(dmin, __shfl_xor_sync(GSR_FULL, dmin, o));
        dmax = max(dmax, __shfl_xor_sync(GSR_FULL, dmax, o));
    }
    if (lane == 0) { red_min[warp] = dmin; red_max[warp] = dmax; }
    if (t == 0) fallback = 0;
    for (uint32_t b = t; b <= B; b += SORT_THREADS) hist[b] = 0;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) { dmin = min(dmin, red_min[w]); dmax = max(dmax, red_max[w]); }
    const float zmin = __uint_as_float(dmin), zmax = __uint_as_float(dmax);  // positive floats: bit order == value order
    const float scale = zmax > zmin ? (float)(B - 1) / (zmax - zmin) : 0.f;
    // ticket = (bucket << 16) | slot-in-bucket, parked in the output array until the scatter (a bucket with
    // >= 65536 keys cannot occur: n <= SORT_CAP)
    for (uint32_t i = t; i < n; i += SORT_THREADS) {
        const float d = __uint_as_float((uint32_t)(g[i] >> 32));
        const uint32_t b = min(B - 1, (uint32_t)((d - zmin) * scale));
        out[i] = (b << 16) | atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
    {   // exclusive scan of hist[0..B) in place, hist[B] = n
        const uint32_t per = B / SORT_THREADS > 0 ? B / SORT_THREADS : 1;  // B is a power of two >= 32
        const uint32_t b0 = t * per;
        uint32_t loc[SORT_BUCKETS / SORT_THREADS];
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < SORT_BUCKETS / SORT_THREADS; k++) {
            loc[k] = 0;
            if (k < per && b0 + k < B) { loc[k] = hist[b0 + k]; }
            sum += loc[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(GSR_FULL, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t base = incl - sum;
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; w++) base += (w < (int)warp) ? wsum[w] : 0u;
#pragma unroll
        for (uint32_t k = 0; k < SORT_BUCKETS / SORT_THREADS; k++) {
            if (k < per && b0 + k < B) { hist[b0 + k] = base; base += loc[k]; }
        }
        if (t == 0) hist[B] = n;
    }
    __syncthreads();
    for (uint32_t i = t; i < n; i += SORT_THREADS) {
        const uint32_t tk = out[i];
        s[hist[tk >> 16] + (tk & 0xffffu)] = g[i];
    }
    __syncthreads();
    for (uint32_t b = t; b < B; b += SORT_THREADS) {  // order the few keys that share a bucket
        const uint32_t lo = hist[b], c = hist[b + 1] - lo;
        if (c > (uint32_t)SORT_BUCKET_MAX) fallback = 1;
        else if (c > 1) {
            for (uint32_t i = 1; i < c; i++) {
                const unsigned long long x = s[lo + i];
                uint32_t j = i;
                while (j > 0 && s[lo + j - 1] > x) { s[lo + j] = s[lo + j - 1]; j--; }
                s[lo + j] = x;
            }
        }
    }
    __syncthreads();
    if (fallback) sort_smem(s, n);  // ends with a barrier
    for (uint32_t i = t; i < n; i += SORT_THREADS) {
        const unsigned long long x = s[i];
        out[i] = (uint32_t)x;
        if (gkeep) gkeep[i] = x;
    }
}

// ---- per-footprint survivor lists ---------------------------------------------------------------------------
// After the sort, every entry of the tile is tested ONCE against the tile's eight 8x4-pixel warp footprints (one thread
// per entry, tile_foot_mask) and the survivors of footprint f are written, in list order, to a compact list that the
// blend warp owning that footprint walks on its own — the blend kernel has no block-level staging, barrier or cull left.
// Lists are carved from one global cursor (counters->foot_total), one atomic per tile; foot_ranges[tile*8+f] = {start, count}.
// A list entry is the Gaussian id, or (store_pos, when a backward pass follows) the absolute position in point_list.
struct FootArgs {
    const float4* records;
    uint32_t* foot_list;
    uint2* foot_ranges;
    gsr_counters* counters;
    uint32_t foot_cap;
    int store_pos;
    int gx;
};
// keys: the n sorted (depth bits << 32 | id) of the tile (shared or global memory); mask8: n bytes of scratch (shared
// memory), or nullptr to park the masks in park32[] (global, 4 bytes per entry — the large-tile path).
__device__ void foot_lists(const FootArgs& fa, const unsigned long long* keys, uint32_t n, uint32_t list_base, int tile,
                           uint8_t* mask8, uint32_t* park32) {
    __shared__ uint32_t f_cnt[GSR_FOOTS], f_off[GSR_FOOTS];
    __shared__ int f_ok;
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ty = tile / fa.gx, tx = tile - ty * fa.gx;
    if (t < GSR_FOOTS) f_cnt[t] = 0;
    __syncthreads();
    uint32_t acc = 0;  // lane f (< 8) of every warp counts footprint f
    for (uint32_t base = warp * 32; base < n; base += SORT_THREADS) {
        const uint32_t i = base + lane;
        uint32_t m = 0;
        if (i < n) {
            const uint32_t id = (uint32_t)keys[i];
            const float4 r0 = fa.records[3 * (size_t)id], r1 = fa.records[3 * (size_t)id + 1];
            m = tile_foot_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r1.w, tx, ty);
            if (mask8) mask8[i] = (uint8_t)m; else park32[i] = m;
        }
#pragma unroll
        for (int f = 0; f < GSR_FOOTS; f++) {
            const uint32_t c = __popc(__ballot_sync(GSR_FULL, (m >> f) & 1u));
            if (lane == (uint32_t)f) acc += c;
        }
    }
    if (lane < GSR_FOOTS && acc) atomicAdd(&f_cnt[lane], acc);
    __syncthreads();
    if (t == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int f = 0; f < GSR_FOOTS; f++) total += f_cnt[f];
        uint32_t start = 0;
        int ok = 1;
        if (total) {
            start = atomicAdd(&fa.counters->foot_total, total);
            if (start + total > fa.foot_cap || start + total < start) { ok = 0; atomicExch(&fa.counters->overflow, 1u); }
        }
        f_ok = ok;
        uint2* fr = fa.foot_ranges + (size_t)tile * GSR_FOOTS;
#pragma unroll
        for (int f = 0; f < GSR_FOOTS; f++) {
            f_off[f] = start;
            fr[f] = ok ? make_uint2(start, f_cnt[f]) : make_uint2(0u, 0u);
            start += f_cnt[f];
        }
    }
    __syncthreads();
    if (f_ok && warp < GSR_FOOTS && f_cnt[warp]) {  // warp f compacts footprint f (SORT_THREADS / 32 == GSR_FOOTS)
        uint32_t run = f_off[warp];
        const uint32_t lt = (1u << lane) - 1u;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t m = i < n ? (mask8 ? (uint32_t)mask8[i] : park32[i]) : 0u;
            const bool keep = (m >> warp) & 1u;
            const uint32_t bal = __ballot_sync(GSR_FULL, keep);
            if (keep) fa.foot_list[run + __popc(bal & lt)] = fa.store_pos ? list_base + i : (uint32_t)keys[i];
            run += __popc(bal);
        }
    }
}
static_assert(SORT_THREADS / 32 == GSR_FOOTS, "one sort warp per footprint");

// Sorts the bucket of one tile (pairs[rg.x..rg.y) by (depth bits, id)) and writes the ids to point_list.
// s: SORT_CAP u64 of shared memory, hist: SORT_BUCKETS+1 u32.  Block-wide (SORT_THREADS threads), ends without a barrier.
__device__ void sort_tile(const uint2 rg, unsigned long long* __restrict__ pairs, uint32_t* __restrict__ point_list, int keep_pairs,
                          unsigned long long* s, uint32_t* hist, const FootArgs& fa, int tile) {
    const uint32_t n = rg.y - rg.x;
    const uint32_t tid = threadIdx.x;
    if (n == 0) {
        if (tid < GSR_FOOTS) fa.foot_ranges[(size_t)tile * GSR_FOOTS + tid] = make_uint2(0u, 0u);
        return;
    }
    unsigned long long* g = pairs + rg.x;
    uint32_t* out = point_list + rg.x;
    if (n <= SORT_CAP) {
        sort_bucket(g, n, out, keep_pairs ? g : nullptr, s, hist);
        __syncthreads();  // the histogram is dead: its storage holds the footprint masks
        foot_lists(fa, s, n, rg.x, tile, reinterpret_cast<uint8_t*>(hist), nullptr);
        return;
    }
    // ---- large tile: chunks sorted in shared memory, cross-chunk steps in global (L2) memory ----
    const uint32_t N = next_pow2(n);
    for (uint32_t c0 = 0; c0 < n; c0 += SORT_CAP) {
        const uint32_t m = min((uint32_t)SORT_CAP, n - c0);
        for (uint32_t i = tid; i < m; i += SORT_THREADS) s[i] = g[c0 + i];
        __syncthreads();
        sort_smem(s, m);
        for (uint32_t i = tid; i < m; i += SORT_THREADS) g[c0 + i] = s[i];
        __syncthreads();
    }
    for (uint32_t k = 2 * SORT_CAP; k <= N; k <<= 1) {
        step_flip(g, n, N, k);
        __syncthreads();
        uint32_t j = k >> 2;
        for (; j >= SORT_CAP; j >>= 1) {
            step_j(g, n, N, j);
            __syncthreads();
        }
        for (uint32_t c0 = 0; c0 < n; c0 += SORT_CAP) {
            const uint32_t m = min((uint32_t)SORT_CAP, n - c0);
            for (uint32_t i = tid; i < m; i += SORT_THREADS) s[i] = g[c0 + i];
            __syncthreads();
            for (uint32_t jj = SORT_CAP >> 1; jj > 0; jj >>= 1) {
                step_j(s, m, SORT_CAP, jj);
                __syncthreads();
            }
            for (uint32_t i = tid; i < m; i += SORT_THREADS) g[c0 + i] = s[i];
            __syncthreads();
        }
    }
    __syncthreads();
    foot_lists(fa, g, n, rg.x, tile, nullptr, out);  // masks parked in point_list until the ids are written
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SORT_THREADS) out[i] = (uint32_t)g[i];
}

// stand-alone per-tile sort kernel (fusing it into the blend prologue was measured and dropped, profiles/r01_experiments.md)
__global__ void __launch_bounds__(SORT_THREADS) k_sort_tiles(const uint2* __restrict__ ranges, unsigned long long* __restrict__ pairs,
                                                             uint32_t* __restrict__ point_list,
                                                             const gsr_counters* __restrict__ counters, int keep_pairs, const FootArgs fa) {
    if (counters->overflow) return;  // set by k_tile_scan, before this kernel started
    __shared__ unsigned long long s[SORT_CAP];
    __shared__ uint32_t hist[SORT_BUCKETS + 1];
    sort_tile(ranges[blockIdx.x], pairs, point_list, keep_pairs, s, hist, fa, (int)blockIdx.x);
}

// =====================================================================================================
// Kernel 5: per-tile front-to-back alpha blend (forward.cu:261-378)
// One CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel footprint.
//   stage   : 256 list entries per batch -> 48-byte records in shared memory (registers prefetch the next batch)
//   cull    : one splat per lane against the warp's footprint (footprint_may_touch: exact box minimum of the quadratic form), ballot
//   compact : surviving records are copied, in order, into the warp's private queue (warp prefix via popc)
//   blend   : the queue is walked by all 32 lanes with the reference's per-pixel arithmetic
// =====================================================================================================
constexpr int BLEND_THREADS = 256;
// NX = number of extra colour channels blended with the same weights (0, or 3 for the product frame's second image)
template <int NX>
struct BlendCfg {
    static constexpr int REC = 48;               // staged record bytes per splat (stride 48 B: conflict-free 128-bit accesses)
    static constexpr int XREC = NX ? 16 : 0;     // staged extra-colour bytes per splat, kept in a separate array (a 64-byte
                                                 // combined stride costs 4-way bank conflicts on every staged load/store)
    static constexpr int PAIR = NX ? 112 : 96;   // queue bytes per splat pair
    static constexpr int QCAP = 62;              // queue entries per warp, an even number (flushed when fewer than 32 slots remain)
    static constexpr int REC_BYTES = 2 * BLEND_THREADS * REC;                  // two staged batches
    static constexpr int XREC_BYTES = 2 * BLEND_THREADS * XREC;
    static constexpr int Q_BYTES = (BLEND_THREADS / 32) * (QCAP / 2) * PAIR;   // per-warp survivor queues
    static constexpr int SMEM = REC_BYTES + XREC_BYTES + Q_BYTES;              // dynamic shared memory of k_blend<NX, .>
};
static_assert(BlendCfg<0>::SMEM <= 48 * 1024, "k_blend<.,0> must fit the default dynamic shared memory limit");

// NX  : extra colour channels `extra[P,NX]` accumulated with the same per-splat weights into `out_extra[NX,H,W]`
//       (+ T_final * bg like the colour image) — what a second rasterizer pass with colors_precomp = extra would
//       return (gaussian_renderer/__init__.py:151-185), without re-running projection, binning, sort and the alpha math.
// NC  : also record n_contrib (the 1-based list position of the last blended splat) for the backward pass.
template <int NX, bool NC>
__global__ void __launch_bounds__(BLEND_THREADS, NX ? 0 : 4) k_blend(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                         const float4* __restrict__ records, const float* __restrict__ extra,
                                                         int W, int H, int gx, const float* __restrict__ bg,
                                                         float* __restrict__ out_color, float* __restrict__ out_depth,
                                                         float* __restrict__ out_alpha, float* __restrict__ out_extra,
                                                         uint32_t* __restrict__ n_contrib,
                                                         const gsr_counters* __restrict__ counters) {
    typedef BlendCfg<NX> Cfg;
    constexpr int REC = Cfg::REC, PAIR = Cfg::PAIR, BLEND_QCAP = Cfg::QCAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sRec = reinterpret_cast<float4*>(smem_raw);
    float4* sQ = reinterpret_cast<float4*>(smem_raw + Cfg::REC_BYTES + Cfg::XREC_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int X0 = blockIdx.x * GSR_TILE + (warp & 1) * 8, Y0 = blockIdx.y * GSR_TILE + (warp >> 1) * 4;
    const int pxi = X0 + (lane & 7), pyi = Y0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    // warp-uniform values are routed through a broadcast so that the compiler keeps them in uniform registers instead of
    // re-deriving them from the thread / CTA ids inside the cull loop (it rematerialises them there to stay at 64 registers)
    const float cx = __shfl_sync(GSR_FULL, (float)X0 + FOOT_HX, 0), cy = __shfl_sync(GSR_FULL, (float)Y0 + FOOT_HY, 0);  // footprint centre
    const uint32_t rec_base = __shfl_sync(GSR_FULL, (uint32_t)__cvta_generic_to_shared(sRec), 0);
    const uint32_t q_base = __shfl_sync(GSR_FULL, (uint32_t)__cvta_generic_to_shared(sQ) + (uint32_t)warp * ((BLEND_QCAP / 2) * PAIR), 0);
    const unsigned lt_mask = (1u << lane) - 1u;

    uint2 range = ranges[tile];
    if (counters->overflow) range = make_uint2(0u, 0u);
    const int n = (int)(range.y - range.x);
    const int nb = (n + BLEND_THREADS - 1) / BLEND_THREADS;

    // T is the running transmittance while the pixel is live.  When the pixel terminates (forward.cu:349-354) T flips
    // its sign: the magnitude keeps the final transmittance, and every later splat fails the `T(1-a) < 1e-4` test on its own
    // (the product is negative), so there is no per-iteration "done" branch.  Pixels outside the image start dead.
    float T = inside ? 1.0f : -1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, E0 = 0.f, E1 = 0.f, E2 = 0.f;
    uint32_t last = 0;
    int qn = 0;  // entries in this warp's queue (warp-uniform)

    // Blend every queued splat into this lane's pixel (reference arithmetic, forward.cu:330-366), two splats per
    // iteration.  The queue stores splats pair-interleaved ({x0,x1},{y0,y1},{a0,a1},{-b0,-b1},{c0,c1},{o0,o1},{r0,r1},
    // {g0,g1},{b0,b1},{depth0,depth1},{pos0,pos1}[,{e0,e0'},{e1,e1'},{e2,e2'}], 96 [112] B per pair) so that the per-splat arithmetic that does not depend
    // on the running transmittance — power, alpha, 1-alpha, colour*alpha — runs on both halves of packed fp32
    // registers (FFMA2/FMUL2/FADD2: two IEEE-rn results per instruction, bit-identical to the scalar ops; the sign of b
    // is folded into the stored -b so the reference's `... - b*dx*dy` needs no negation).  Only expf, the 0.99 clamp
    // and the serial transmittance updates stay scalar.  Branch-free: the three skip rules (power > 0, alpha < 1/255,
    // T(1-alpha) < 1e-4) are predicates and a skipped splat contributes through a weight of exactly 0.
    const f32x2 npx2 = pk2(-pixx, -pixx), npy2 = pk2(-pixy, -pixy), mhalf2 = pk2(-0.5f, -0.5f), mone2 = pk2(-1.0f, -1.0f),
                one2 = pk2(1.0f, 1.0f);
    auto drain = [&]() {
        if (qn & 1) {  // complete the last pair with a splat that can never hit (opacity 0)
            if (lane < 11 + NX) sts32(q_base + (uint32_t)(qn >> 1) * PAIR + 4 + lane * 8, 0.0f);
        }
        __syncwarp();
        uint32_t qa = q_base;
        const int np = (qn + 1) >> 1;
        for (int k = 0; k < np; k++, qa += PAIR) {
            const float4 L0 = lds128(qa), L1 = lds128(qa + 16), L2 = lds128(qa + 32), L3 = lds128(qa + 48), L4 = lds128(qa + 64);
            float4 L5, L6;  // {pos0,pos1[,e0,e0']}, {e1,e1',e2,e2'}
            if (NX) { L5 = lds128(qa + 80); L6 = lds128(qa + 96); }
            else { const float2 t = lds64(qa + 80); L5 = make_float4(t.x, t.y, 0.f, 0.f); L6 = L5; }
            const f32x2 dx = add2(pk2(L0.x, L0.y), npx2), dy = add2(pk2(L0.z, L0.w), npy2);
            const f32x2 t1 = mul2(pk2(L2.x, L2.y), dy);   // c * dy
            const f32x2 t3 = mul2(pk2(L1.x, L1.y), dx);   // a * dx
            const f32x2 t2 = mul2(pk2(L1.z, L1.w), dx);   // (-b) * dx
            const f32x2 t4 = mul2(dy, t1);                // dy * (c dy)
            const f32x2 t5 = mul2(dy, t2);                // -(dy * (b dx))
            const f32x2 t6 = fma2(dx, t3, t4);            // a dx^2 + c dy^2
            const f32x2 pw = fma2(t6, mhalf2, t5);        // power = -0.5 (a dx^2 + c dy^2) - b dx dy
            float p0, p1;
            upk2(pw, p0, p1);
            float a0, a1;
            upk2(mul2(pk2(L2.z, L2.w), pk2(exp(p0), exp(p1))), a0, a1);  // opacity * exp(power); a packed expf was measured slower
            a0 = min(0.99f, a0);
            a1 = min(0.99f, a1);
            const bool hit0 = !(p0 > 0.0f) && !(a0 < 1.0f / 255.0f), hit1 = !(p1 > 0.0f) && !(a1 < 1.0f / 255.0f);
            const f32x2 al = pk2(a0, a1);
            float om0, om1;
            upk2(fma2(al, mone2, one2), om0, om1);        // 1 - alpha
            float cr0, cr1, cg0, cg1, cb0, cb1, cd0, cd1;
            upk2(mul2(pk2(L3.x, L3.y), al), cr0, cr1);    // colour * alpha
            upk2(mul2(pk2(L3.z, L3.w), al), cg0, cg1);
            upk2(mul2(pk2(L4.x, L4.y), al), cb0, cb1);
            upk2(mul2(pk2(L4.z, L4.w), al), cd0, cd1);
            float ex0 = 0.f, ex1 = 0.f, ey0 = 0.f, ey1 = 0.f, ez0 = 0.f, ez1 = 0.f;
            if (NX) {
                upk2(mul2(pk2(L5.z, L5.w), al), ex0, ex1);
                upk2(mul2(pk2(L6.x, L6.y), al), ey0, ey1);
                upk2(mul2(pk2(L6.z, L6.w), al), ez0, ez1);
            }
            {   // first splat of the pair
                const bool act = hit0 && T > 0.0f;              // the pixel is live and the splat is not skipped
                const float test_T = T * om0;
                const bool live = act && !(test_T < 0.0001f);   // ... and it does not terminate the pixel: blend it
                const float Tw = live ? T : 0.0f;
                C0 = fmaf(Tw, cr0, C0); C1 = fmaf(Tw, cg0, C1); C2 = fmaf(Tw, cb0, C2); Dp = fmaf(Tw, cd0, Dp);
                if (NX) { E0 = fmaf(Tw, ex0, E0); E1 = fmaf(Tw, ey0, E1); E2 = fmaf(Tw, ez0, E2); }
                T = act ? (live ? test_T : -T) : T;             // blended / terminated (sign flip) / untouched
                if (NC) last = live ? __float_as_uint(L5.x) : last;
            }
            {   // second splat of the pair
                const bool act = hit1 && T > 0.0f;
                const float test_T = T * om1;
                const bool live = act && !(test_T < 0.0001f);
                const float Tw = live ? T : 0.0f;
                C0 = fmaf(Tw, cr1, C0); C1 = fmaf(Tw, cg1, C1); C2 = fmaf(Tw, cb1, C2); Dp = fmaf(Tw, cd1, Dp);
                if (NX) { E0 = fmaf(Tw, ex1, E0); E1 = fmaf(Tw, ey1, E1); E2 = fmaf(Tw, ez1, E2); }
                T = act ? (live ? test_T : -T) : T;
                if (NC) last = live ? __float_as_uint(L5.y) : last;
            }
        }
        qn = 0;
        __syncwarp();
    };

    // Software pipeline over batches of 256 list entries, double-buffered in shared memory with ONE barrier per batch:
    // while batch b is culled/blended out of buffer b&1, the records of batch b+1 (already in registers, gathered
    // during batch b-1) are stored into the other buffer and the gather of batch b+2 is issued.
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra, rd = ra;
    uint32_t id_next = 0;
    auto gather = [&](int batch, uint32_t id) {  // records of list entry (batch, tid) -> registers
        if (batch * BLEND_THREADS + tid < n) {
            const float4* r = records + 3 * (size_t)id;
            ra = r[0]; rb = r[1]; rc = r[2];
            rc.w = __uint_as_float((uint32_t)(batch * BLEND_THREADS + tid + 1));  // 1-based position in the tile list
            if (NX) {
                const float* e = extra + 3 * (size_t)id;
                rd = make_float4(e[0], e[1], e[2], 0.0f);
            }
        }
    };
    auto stage = [&](int batch) {  // registers -> buffer batch&1
        if (batch * BLEND_THREADS + tid < n) {
            const uint32_t sa = rec_base + (uint32_t)((batch & 1) * BLEND_THREADS + tid) * REC;
            sts128(sa, ra); sts128(sa + 16, rb); sts128(sa + 32, rc);
            if (NX) sts128(rec_base + Cfg::REC_BYTES + (uint32_t)((batch & 1) * BLEND_THREADS + tid) * 16, rd);
        }
    };
    if (tid < n) gather(0, point_list[range.x + tid]);
    if (BLEND_THREADS + tid < n) id_next = point_list[range.x + BLEND_THREADS + tid];
    stage(0);
    gather(1, id_next);
    if (2 * BLEND_THREADS + tid < n) id_next = point_list[range.x + 2 * BLEND_THREADS + tid];
    __syncthreads();
    bool warp_done = false;

    for (int b = 0; b < nb; b++) {
        const int cnt = min(BLEND_THREADS, n - b * BLEND_THREADS);
        const uint32_t buf = rec_base + (uint32_t)((b & 1) * BLEND_THREADS) * REC;
        if (!warp_done) {
            for (int base = 0; base < cnt; base += 32) {
                const int s = base + lane;
                const uint32_t sa = buf + (uint32_t)s * REC;
                bool keep = false;
                float4 A, B;
                if (s < cnt) {
                    A = lds128(sa); B = lds128(sa + 16);
                    keep = footprint_may_touch(A.x - cx, A.y - cy, A.z, A.w, B.x, B.w);
                }
                const unsigned mask = __ballot_sync(GSR_FULL, keep);
                if (mask) {
                    if (keep) {
                        const uint32_t q = (uint32_t)(qn + __popc(mask & lt_mask));
                        const uint32_t qa = q_base + (q >> 1) * PAIR + (q & 1) * 4;
                        const float4 Cc = lds128(sa + 32);
                        sts32(qa, A.x); sts32(qa + 8, A.y); sts32(qa + 16, A.z); sts32(qa + 24, -A.w);
                        sts32(qa + 32, B.x); sts32(qa + 40, B.y); sts32(qa + 48, Cc.x); sts32(qa + 56, Cc.y);
                        sts32(qa + 64, Cc.z); sts32(qa + 72, B.z); sts32(qa + 80, Cc.w);
                        if (NX) {
                            const float4 Dd = lds128(rec_base + Cfg::REC_BYTES + (uint32_t)((b & 1) * BLEND_THREADS + s) * 16);
                            sts32(qa + 88, Dd.x); sts32(qa + 96, Dd.y); sts32(qa + 104, Dd.z);
                        }
                    }
                    qn += __popc(mask);
                    if (qn > BLEND_QCAP - 32) {
                        drain();
                        if (__all_sync(GSR_FULL, T < 0.0f)) { warp_done = true; break; }
                    }
                }
            }
        }
        if (b + 1 < nb) {
            stage(b + 1);
            gather(b + 2, id_next);
            if ((b + 3) * BLEND_THREADS + tid < n) id_next = point_list[range.x + (b + 3) * BLEND_THREADS + tid];
            // whole tile finished?  (also publishes buffer (b+1)&1 and retires buffer b&1)
            if (__syncthreads_count(T < 0.0f) == BLEND_THREADS) break;
        }
    }
    if (qn) drain();
    if (inside) {
        const float T_out = fabsf(T);  // final transmittance, whether the pixel terminated or the list ran out
        const size_t pid = (size_t)W * pyi + pxi;
        const size_t HW = (size_t)H * W;
        out_alpha[pid] = 1 - T_out;
        if (NC) n_contrib[pid] = last;
        out_color[pid] = C0 + T_out * bg[0];
        out_color[HW + pid] = C1 + T_out * bg[1];
        out_color[2 * HW + pid] = C2 + T_out * bg[2];
        out_depth[pid] = Dp;
        if (NX) {
            out_extra[pid] = E0 + T_out * bg[0];
            out_extra[HW + pid] = E1 + T_out * bg[1];
            out_extra[2 * HW + pid] = E2 + T_out * bg[2];
        }
    }
}

// =====================================================================================================
// Second pass over identical geometry (GSR_FLAG_REUSE_GEOMETRY): the product frame renders every camera twice with
// the same Gaussians — SH colours, then colors_precomp = normals (gaussian_renderer/__init__.py:151-185,
// sugar_model.py:2141-2183).  Projection, binning and sorting of the first pass stay valid; only the colour slot of
// the visible records is rewritten before the blend.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_recolor(int P, const int* __restrict__ radii, const float* __restrict__ colors_precomp,
                                                 float4* __restrict__ records) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || radii[idx] <= 0) return;
    float* c = reinterpret_cast<float*>(records + 3 * (size_t)idx + 2);  // .w (log2 opacity) stays
    c[0] = colors_precomp[3 * (size_t)idx];
    c[1] = colors_precomp[3 * (size_t)idx + 1];
    c[2] = colors_precomp[3 * (size_t)idx + 2];
}

// =====================================================================================================
// optional per-kernel timing (bench.py roofline): CUDA events recorded around each forward kernel on the
// launching stream; no effect unless gsr_profile_begin() was called.  Not thread safe (one profiled stream).
// =====================================================================================================
struct ProfileState {
    bool on = false;
    int max_frames = 0, frames = 0;
    int stride = 1, seen = 0;  // every stride-th forward call is timed (the event records cost ~1.5 % of a frame)
    cudaEvent_t* ev = nullptr;  // 6 per frame
    int allocated = 0;
};
static ProfileState g_prof;
static inline void prof_mark(int k, cudaStream_t st) {
    if (g_prof.on && g_prof.frames < g_prof.max_frames && g_prof.seen % g_prof.stride == 0) cudaEventRecord(g_prof.ev[g_prof.frames * 6 + k], st);
}
int profile_begin(int max_frames, int stride) {
    if (max_frames <= 0 || stride <= 0) { set_error("gsr_profile_begin: max_frames and stride must be > 0"); return GSR_ERR_INVALID; }
    if (g_prof.allocated < max_frames * 6) {
        cudaEvent_t* ne = new cudaEvent_t[max_frames * 6];
        for (int i = 0; i < max_frames * 6; i++) {
            if (i < g_prof.allocated) ne[i] = g_prof.ev[i];
            else if (cudaEventCreate(&ne[i]) != cudaSuccess) { set_error("gsr_profile_begin: cudaEventCreate failed"); return GSR_ERR_CUDA; }
        }
        delete[] g_prof.ev;
        g_prof.ev = ne;
        g_prof.allocated = max_frames * 6;
    }
    g_prof.max_frames = max_frames;
    g_prof.frames = 0;
    g_prof.stride = stride;
    g_prof.seen = 0;
    g_prof.on = true;
    return GSR_OK;
}
int profile_end(float* ms, int* frames) {
    g_prof.on = false;
    double acc[5] = {0, 0, 0, 0, 0};
    for (int f = 0; f < g_prof.frames; f++) {
        if (cudaEventSynchronize(g_prof.ev[f * 6 + 5]) != cudaSuccess) { set_error("gsr_profile_end: event sync failed"); return GSR_ERR_CUDA; }
        for (int k = 0; k < 5; k++) {
            float t = 0;
            cudaEventElapsedTime(&t, g_prof.ev[f * 6 + k], g_prof.ev[f * 6 + k + 1]);
            acc[k] += t;
        }
    }
    for (int k = 0; k < 5; k++) ms[k] = g_prof.frames ? (float)(acc[k] / g_prof.frames) : 0.f;
    if (frames) *frames = g_prof.frames;
    return GSR_OK;
}

// =====================================================================================================
// host side
// =====================================================================================================
static int sh_bulk_mode() {  // GSR_SH_STAGING=cpasync selects the LDGSTS path, default is the TMA bulk copy
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("GSR_SH_STAGING");
        mode = (e && strcmp(e, "cpasync") == 0) ? 0 : 1;
    }
    return mode;
}
template <int DEG>
static void launch_pre(bool vec, bool win, const PreParams& pp, cudaStream_t st) {
    const int grid = (pp.P + PRE_THREADS - 1) / PRE_THREADS;
    if (win && sh_bulk_mode()) k_preprocess<DEG, true, true, true><<<grid, PRE_THREADS, 0, st>>>(pp);
    else if (vec && sh_bulk_mode()) k_preprocess<DEG, true, true><<<grid, PRE_THREADS, 0, st>>>(pp);
    else if (vec) k_preprocess<DEG, true, false><<<grid, PRE_THREADS, 0, st>>>(pp);
    else k_preprocess<DEG, false, false><<<grid, PRE_THREADS, 0, st>>>(pp);
}

template <int NX, bool NC>
static void launch_blend_t(const BlendArgs& a, cudaStream_t st) {
    typedef BlendCfg<NX> Cfg;
    if (NX) {  // > 48 KB of dynamic shared memory: opt in once per device
        static bool configured[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !configured[dev]) {
            cudaFuncSetAttribute(k_blend<NX, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
            configured[dev] = true;
        }
    }
    k_blend<NX, NC><<<dim3(a.gx, a.gy), BLEND_THREADS, Cfg::SMEM, st>>>(a.ranges, a.point_list, a.records, a.extra, a.W, a.H, a.gx, a.bg,
                                                                       a.out_color, a.out_depth, a.out_alpha, a.out_extra, a.n_contrib, a.counters);
}
// The 6-channel variant runs at 70 registers / 3 CTAs per SM; holding it to 64 registers / 4 CTAs (shorter queues, 92 B of
// spills) was measured slower: 780 vs 862 product frames/s (profiles/r01_experiments.md).
static int blend_legacy() {  // GSR_BLEND=legacy: the round-1 tile-staged blend (bit-exact only), kept for A/B measurements
    static int m = -1;
    if (m < 0) { const char* e = getenv("GSR_BLEND"); m = (e && strcmp(e, "legacy") == 0) ? 1 : 0; }
    return m;
}
static void launch_blend(const BlendArgs& a, cudaStream_t st) {
    if (!blend_legacy()) { launch_blend_lists(a, st); return; }
    if (a.extra) { if (a.n_contrib) launch_blend_t<3, true>(a, st); else launch_blend_t<3, false>(a, st); }
    else         { if (a.n_contrib) launch_blend_t<0, true>(a, st); else launch_blend_t<0, false>(a, st); }
}

int forward_impl(const gsr_frame* f, const gsr_workspace* ws, float* out_color, float* out_depth, float* out_alpha,
                 int32_t* radii, const float* extra_colors, float* out_extra, int flags, cudaStream_t st) {
    if (!f || !ws) { set_error("gsr_forward: null frame/workspace"); return GSR_ERR_INVALID; }
    if (f->P < 0 || f->W <= 0 || f->H <= 0) { set_error("gsr_forward: bad sizes P=%d W=%d H=%d", f->P, f->W, f->H); return GSR_ERR_INVALID; }
    if (!out_color || !out_depth || !out_alpha) { set_error("gsr_forward: null output image"); return GSR_ERR_INVALID; }
    if ((extra_colors == nullptr) != (out_extra == nullptr)) { set_error("gsr_forward_multi: extra_colors and out_extra go together"); return GSR_ERR_INVALID; }
    const size_t HW = (size_t)f->W * f->H;
    const ImageLayout il(f->W, f->H);
    if (!ws->image || ws->image_bytes < il.total) { set_error("gsr_forward: image workspace too small (%zu < %zu)", ws->image_bytes, il.total); return GSR_ERR_WORKSPACE; }
    char* img = (char*)ws->image;
    gsr_counters* counters = (gsr_counters*)(img + il.counters);
    const bool debug = f->debug != 0;
    if (f->P == 0) {  // rasterize_points.cu:68-71,82: zero images, nothing else runs
        cudaMemsetAsync(out_color, 0, 12 * HW, st);
        cudaMemsetAsync(out_depth, 0, 4 * HW, st);
        cudaMemsetAsync(out_alpha, 0, 4 * HW, st);
        if (out_extra) cudaMemsetAsync(out_extra, 0, 12 * HW, st);
        cudaMemsetAsync(img, 0, il.zero_bytes(), st);
        cudaMemsetAsync(img + il.ranges, 0, 8 * (size_t)il.tiles, st);
        return check_launch("gsr_forward(P=0)", debug, st);
    }
    if (!radii || !f->means3D || !f->opacities || !f->viewmatrix || !f->projmatrix || !f->campos || !f->bg) {
        set_error("gsr_forward: null required input");
        return GSR_ERR_INVALID;
    }
    if ((f->shs == nullptr) == (f->colors_precomp == nullptr)) { set_error("gsr_forward: provide exactly one of shs / colors_precomp"); return GSR_ERR_INVALID; }
    const bool has_sr = f->scales != nullptr && f->rotations != nullptr;
    if (has_sr == (f->cov3D_precomp != nullptr) || ((f->scales != nullptr) != (f->rotations != nullptr))) {
        set_error("gsr_forward: provide exactly one of scales+rotations / cov3D_precomp");
        return GSR_ERR_INVALID;
    }
    const int D = f->D < 0 ? 0 : (f->D > 3 ? 3 : f->D);
    if (f->shs && (D + 1) * (D + 1) > f->M) { set_error("gsr_forward: sh degree %d needs %d coefficients, shs has M=%d", D, (D + 1) * (D + 1), f->M); return GSR_ERR_INVALID; }
    const GeomLayout gl((size_t)f->P);
    if (!ws->geom || ws->geom_bytes < gl.total) { set_error("gsr_forward: geometry workspace too small (%zu < %zu)", ws->geom_bytes, gl.total); return GSR_ERR_WORKSPACE; }
    const size_t cap = ws->binning ? BinLayout::capacity_of(ws->binning_bytes) : 0;
    if (cap < 1) { set_error("gsr_forward: binning workspace too small"); return GSR_ERR_WORKSPACE; }
    const BinLayout bl(cap);
    char* geo = (char*)ws->geom;
    char* bin = (char*)ws->binning;

    if (flags & GSR_FLAG_REUSE_GEOMETRY) {
        // `radii` is an INPUT here: the radii of the pass whose workspaces are being reused
        if (!f->colors_precomp) { set_error("gsr_forward: GSR_FLAG_REUSE_GEOMETRY needs colors_precomp"); return GSR_ERR_INVALID; }
        if (flags & GSR_FLAG_FOR_BACKWARD) { set_error("gsr_forward: GSR_FLAG_REUSE_GEOMETRY cannot be combined with GSR_FLAG_FOR_BACKWARD"); return GSR_ERR_INVALID; }
        k_recolor<<<(f->P + 255) / 256, 256, 0, st>>>(f->P, radii, f->colors_precomp, (float4*)(geo + gl.records));
        int rc0 = check_launch("gsr_forward/recolor", debug, st);
        if (rc0) return rc0;
        BlendArgs ba{(const uint2*)(img + il.ranges), (const uint32_t*)(bin + bl.point_list), (const float4*)(geo + gl.records), extra_colors,
                     f->W, f->H, il.gx, il.gy, f->bg, out_color, out_depth, out_alpha, out_extra, nullptr, counters,
                     (const uint2*)(img + il.foot_ranges), (const uint32_t*)(bin + bl.foot_list), (flags & GSR_FLAG_EXACT_IMAGES) ? 1 : 0};
        launch_blend(ba, st);
        return check_launch("gsr_forward/blend(reuse)", debug, st);
    }

    cudaMemsetAsync(img, 0, il.zero_bytes(), st);
    prof_mark(0, st);

    PreParams pp;
    pp.P = f->P; pp.D = D; pp.M = f->M; pp.W = f->W; pp.H = f->H; pp.gx = il.gx; pp.gy = il.gy;
    pp.scale_modifier = f->scale_modifier; pp.tanfovx = f->tanfovx; pp.tanfovy = f->tanfovy;
    pp.focal_y = f->H / (2.0f * f->tanfovy);  // rasterizer_impl.cu:223-224
    pp.focal_x = f->W / (2.0f * f->tanfovx);
    pp.rot_vec = (((uintptr_t)f->rotations & 15) == 0) ? 1 : 0;
    pp.tight = (flags & GSR_FLAG_TIGHT_TILES) ? 1 : 0;
    pp.prefiltered = f->prefiltered; pp.for_backward = (flags & GSR_FLAG_FOR_BACKWARD) ? 1 : 0;
    pp.means3D = f->means3D; pp.shs = f->shs; pp.colors_precomp = f->colors_precomp; pp.opacities = f->opacities;
    pp.scales = f->scales; pp.rotations = f->rotations; pp.cov3D_precomp = f->cov3D_precomp;
    pp.view = f->viewmatrix; pp.proj = f->projmatrix; pp.campos = f->campos;
    pp.records = (float4*)(geo + gl.records); pp.cov3D = (float*)(geo + gl.cov3D); pp.clamped = (uint8_t*)(geo + gl.clamped);
    pp.radii = radii; pp.tile_count = (uint32_t*)(img + il.tile_count); pp.tile_big = (uint32_t*)(img + il.tile_big);
    pp.ranks = (uint32_t*)(geo + gl.ranks); pp.counters = counters;

    if (f->colors_precomp) launch_pre<-1>(false, false, pp, st);
    else {
        const bool vec = (f->M % 4 == 0) && (((uintptr_t)f->shs & 15) == 0);
        // 4-byte aligned rows (e.g. M = 25): the aligned window of sh_nv(D, true) float4 must fit inside every row
        const bool win = !vec && (((uintptr_t)f->shs & 15) == 0) && (size_t)f->M * 12 >= (size_t)sh_nv(D, true) * 16;
        switch (D) {
            case 0: launch_pre<0>(vec, win, pp, st); break;
            case 1: launch_pre<1>(vec, win, pp, st); break;
            case 2: launch_pre<2>(vec, win, pp, st); break;
            default: launch_pre<3>(vec, win, pp, st); break;
        }
    }
    prof_mark(1, st);
    int rc = check_launch("gsr_forward/preprocess", debug, st);
    if (rc) return rc;

    uint2* ranges = (uint2*)(img + il.ranges);
    k_tile_scan<<<1, 1024, 0, st>>>(pp.tile_count, pp.tile_big, (uint32_t*)(img + il.tile_fill), ranges, counters, il.tiles, (uint32_t)(cap > 0xffffffffull ? 0xffffffffull : cap));
    prof_mark(2, st);
    if ((rc = check_launch("gsr_forward/tile_scan", debug, st))) return rc;

    if (pp.tight)
        k_emit<true><<<(f->P + 255) / 256, 256, 0, st>>>(f->P, il.gx, il.gy, radii, pp.records, pp.ranks, ranges, (uint32_t*)(img + il.tile_fill),
                                                         (uint2*)(bin + bl.pairs), counters);
    else
        k_emit<false><<<(f->P + 255) / 256, 256, 0, st>>>(f->P, il.gx, il.gy, radii, pp.records, pp.ranks, ranges, (uint32_t*)(img + il.tile_fill),
                                                          (uint2*)(bin + bl.pairs), counters);
    prof_mark(3, st);
    if ((rc = check_launch("gsr_forward/emit", debug, st))) return rc;

    const int keep_pairs = (flags & GSR_FLAG_SORTED_KEYS) ? 1 : 0;
    uint32_t* n_contrib = (flags & GSR_FLAG_FOR_BACKWARD) ? (uint32_t*)(img + il.n_contrib) : nullptr;
    FootArgs fa{pp.records, (uint32_t*)(bin + bl.foot_list), (uint2*)(img + il.foot_ranges), counters,
                (uint32_t)(bl.foot_capacity > 0xffffffffull ? 0xffffffffull : bl.foot_capacity), n_contrib ? 1 : 0, il.gx};
    k_sort_tiles<<<il.tiles, SORT_THREADS, 0, st>>>(ranges, (unsigned long long*)(bin + bl.pairs), (uint32_t*)(bin + bl.point_list), counters, keep_pairs, fa);
    prof_mark(4, st);
    if ((rc = check_launch("gsr_forward/sort", debug, st))) return rc;
    BlendArgs ba{ranges, (const uint32_t*)(bin + bl.point_list), pp.records, extra_colors, f->W, f->H, il.gx, il.gy, f->bg,
                 out_color, out_depth, out_alpha, out_extra, n_contrib, counters,
                 (const uint2*)(img + il.foot_ranges), (const uint32_t*)(bin + bl.foot_list), (flags & GSR_FLAG_EXACT_IMAGES) ? 1 : 0};
    launch_blend(ba, st);
    prof_mark(5, st);
    if (g_prof.on) {
        if (g_prof.frames < g_prof.max_frames && g_prof.seen % g_prof.stride == 0) g_prof.frames++;
        g_prof.seen++;
    }
    return check_launch("gsr_forward/blend", debug, st);
}

}  // namespace gsr

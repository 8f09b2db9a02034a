// gsr_b200 — distCUDA2: mean squared distance to the 3 nearest neighbours.
//
// Replaces SimpleKNN::knn (KNN/simple_knn.cu:185-220).  Same algorithm — Morton order over the bounding box
// (which always contains the origin, simple_knn.cu:191), 1024-point boxes, per-point seed from the +-3 Morton
// neighbours, then an exhaustive scan of every box whose AABB distance is within the current bound — so the
// result is the exact 3-NN mean.  Unlike the reference there is no host round trip (its two blocking
// cudaMemcpy's of the bounding box, :197,:200), no cudaMalloc/cudaFree, no Thrust/CUB: the bounding box
// stays on the device, and the (morton, index) sort is a small hand-written stable LSD radix sort
// (one warp per 2048-key chunk, warp match_any ranking).  Init-time only (gaussian_model.py:144).
#include "gsr_common.cuh"
#include <cfloat>

namespace gsr {

constexpr int KNN_BOX = 1024;
constexpr int KNN_CHUNK = 2048;  // keys per warp in the radix sort

struct KnnLayout {
    size_t bbox, codes0, codes1, idx0, idx1, hist, boxes, total;
    int nchunks, nboxes;
    __host__ __device__ explicit KnnLayout(size_t P) {
        nchunks = (int)((P + KNN_CHUNK - 1) / KNN_CHUNK);
        nboxes = (int)((P + KNN_BOX - 1) / KNN_BOX);
        size_t o = 0;
        bbox = o;   o = align_up(o + 32, 256);
        codes0 = o; o = align_up(o + 4 * P, 256);
        codes1 = o; o = align_up(o + 4 * P, 256);
        idx0 = o;   o = align_up(o + 4 * P, 256);
        idx1 = o;   o = align_up(o + 4 * P, 256);
        hist = o;   o = align_up(o + 4 * 256 * (size_t)nchunks, 256);
        boxes = o;  o = align_up(o + 24 * (size_t)nboxes, 256);
        total = o + 256;
    }
};
size_t dist2_bytes(int P) { return KnnLayout((size_t)(P < 0 ? 0 : P)).total; }

__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_bbox_init(uint32_t* bbox) {
    if (threadIdx.x < 6) bbox[threadIdx.x] = f2ord(0.0f);  // reduction init {0,0,0} (simple_knn.cu:191)
}
__global__ void __launch_bounds__(256) k_bbox(int P, const float* __restrict__ pts, uint32_t* bbox) {
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(GSR_FULL, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(GSR_FULL, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&bbox[k], f2ord(mn[k]));
            atomicMax(&bbox[3 + k], f2ord(mx[k]));
        }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {  // simple_knn.cu:46-53
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
__global__ void __launch_bounds__(256) k_morton(int P, const float* __restrict__ pts, const uint32_t* __restrict__ bbox,
                                                uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float mnx = ord2f(bbox[0]), mny = ord2f(bbox[1]), mnz = ord2f(bbox[2]);
    const float mxx = ord2f(bbox[3]), mxy = ord2f(bbox[4]), mxz = ord2f(bbox[5]);
    const float3 c = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    const uint32_t x = prep_morton(((c.x - mnx) / (mxx - mnx)) * ((1 << 10) - 1));  // simple_knn.cu:55-61
    const uint32_t y = prep_morton(((c.y - mny) / (mxy - mny)) * ((1 << 10) - 1));
    const uint32_t z = prep_morton(((c.z - mnz) / (mxz - mnz)) * ((1 << 10) - 1));
    codes[i] = x | (y << 1) | (z << 2);
    idx[i] = (uint32_t)i;
}

// ---- stable LSD radix sort, 8-bit digits, one warp per chunk of KNN_CHUNK consecutive keys ----
constexpr int RS_WARPS = 4;
__global__ void __launch_bounds__(RS_WARPS * 32) k_rs_hist(int P, const uint32_t* __restrict__ keys, int shift, int nchunks,
                                                           uint32_t* __restrict__ hist) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x * RS_WARPS + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (chunk < nchunks) {
        const int b = chunk * KNN_CHUNK, e = min(P, b + KNN_CHUNK);
        for (int i = b + lane; i < e; i += 32) atomicAdd(&cnt[warp][(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(size_t)d * nchunks + chunk] = cnt[warp][d];
    }
}
__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t* hist, int n) {  // exclusive scan of n entries, one CTA
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const uint32_t v = i < n ? hist[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(GSR_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t s = wsum[lane];
            uint32_t si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(GSR_FULL, si, o);
                if (lane >= o) si += t;
            }
            wsum[lane] = si - s;
        }
        __syncthreads();
        const uint32_t excl = carry + wsum[warp] + incl - v;
        if (i < n) hist[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(RS_WARPS * 32) k_rs_scatter(int P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                              int shift, int nchunks, const uint32_t* __restrict__ hist,
                                                              uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t off[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x * RS_WARPS + warp;
    if (chunk >= nchunks) return;
    for (int d = lane; d < 256; d += 32) off[warp][d] = hist[(size_t)d * nchunks + chunk];
    __syncwarp();
    const int b = chunk * KNN_CHUNK, e = min(P, b + KNN_CHUNK);
    for (int i0 = b; i0 < e; i0 += 32) {
        const int i = i0 + lane;
        const bool act = i < e;
        const uint32_t k = act ? keys[i] : 0u, v = act ? vals[i] : 0u;
        const uint32_t d = act ? ((k >> shift) & 255u) : 256u + lane;  // inactive lanes get unique pseudo-digits
        const unsigned peers = __match_any_sync(GSR_FULL, d);
        const int leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t basepos = 0;
        if (act && lane == leader) {
            basepos = off[warp][d];
            off[warp][d] = basepos + __popc(peers);
        }
        basepos = __shfl_sync(GSR_FULL, basepos, leader);
        if (act) {
            keys_out[basepos + rank] = k;
            vals_out[basepos + rank] = v;
        }
        __syncwarp();
    }
}

// ---- boxes ----
__global__ void __launch_bounds__(256) k_box_minmax(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order,
                                                    float* __restrict__ boxes) {
    __shared__ float red[8][6];
    const int b = blockIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = b * KNN_BOX + threadIdx.x; i < min(P, (b + 1) * KNN_BOX); i += 256) {
        const uint32_t id = order[i];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)id + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(GSR_FULL, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(GSR_FULL, mx[k], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { red[warp][k] = mn[k]; red[warp][3 + k] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[0][threadIdx.x];
        for (int w = 1; w < 8; w++) v = threadIdx.x < 3 ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
        boxes[6 * (size_t)b + threadIdx.x] = v;
    }
}

__device__ __forceinline__ void kbest3(const float3& ref, const float3& point, float* knn) {  // simple_knn.cu:130-145
    float3 d = {point.x - ref.x, point.y - ref.y, point.z - ref.z};
    float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (knn[j] > dist) {
            float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
    }
}
__device__ __forceinline__ float3 ldp(const float* pts, uint32_t id) {
    return make_float3(pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2]);
}
__global__ void __launch_bounds__(256) k_box_mean_dist(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order,
                                                       const float* __restrict__ boxes, int nboxes, float* __restrict__ dists) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const uint32_t me = order[idx];
    const float3 point = ldp(pts, me);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); i++) {
        if (i == idx) continue;
        kbest3(point, ldp(pts, order[i]), best);
    }
    const float reject = best[2];
    best[0] = FLT_MAX; best[1] = FLT_MAX; best[2] = FLT_MAX;
    for (int b = 0; b < nboxes; b++) {
        const float* bx = boxes + 6 * (size_t)b;
        float3 diff = {0, 0, 0};  // distBoxPoint, simple_knn.cu:119-128
        if (point.x < bx[0] || point.x > bx[3]) diff.x = fminf(fabsf(point.x - bx[0]), fabsf(point.x - bx[3]));
        if (point.y < bx[1] || point.y > bx[4]) diff.y = fminf(fabsf(point.y - bx[1]), fabsf(point.y - bx[4]));
        if (point.z < bx[2] || point.z > bx[5]) diff.z = fminf(fabsf(point.z - bx[2]), fabsf(point.z - bx[5]));
        const float dist = diff.x * diff.x + diff.y * diff.y + diff.z * diff.z;
        if (dist > reject || dist > best[2]) continue;
        for (int i = b * KNN_BOX; i < min(P, (b + 1) * KNN_BOX); i++) {
            if (i == idx) continue;
            kbest3(point, ldp(pts, order[i]), best);
        }
    }
    dists[me] = (best[0] + best[1] + best[2]) / 3.0f;
}

int dist2_impl(int P, const float* points, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (P < 0) { set_error("gsr_dist2: negative P"); return GSR_ERR_INVALID; }
    if (P == 0) return GSR_OK;
    if (!points || !out) { set_error("gsr_dist2: null pointer"); return GSR_ERR_INVALID; }
    const KnnLayout L((size_t)P);
    if (!ws || ws_bytes < L.total) { set_error("gsr_dist2: workspace too small (%zu < %zu)", ws_bytes, L.total); return GSR_ERR_WORKSPACE; }
    char* w = (char*)ws;
    uint32_t* bbox = (uint32_t*)(w + L.bbox);
    uint32_t *k0 = (uint32_t*)(w + L.codes0), *k1 = (uint32_t*)(w + L.codes1), *v0 = (uint32_t*)(w + L.idx0), *v1 = (uint32_t*)(w + L.idx1);
    uint32_t* hist = (uint32_t*)(w + L.hist);
    float* boxes = (float*)(w + L.boxes);
    k_bbox_init<<<1, 32, 0, st>>>(bbox);
    k_bbox<<<min(148 * 8, (P + 255) / 256), 256, 0, st>>>(P, points, bbox);
    k_morton<<<(P + 255) / 256, 256, 0, st>>>(P, points, bbox, k0, v0);
    const int rs_blocks = (L.nchunks + RS_WARPS - 1) / RS_WARPS;
    for (int pass = 0; pass < 4; pass++) {
        k_rs_hist<<<rs_blocks, RS_WARPS * 32, 0, st>>>(P, k0, pass * 8, L.nchunks, hist);
        k_rs_scan<<<1, 1024, 0, st>>>(hist, 256 * L.nchunks);
        k_rs_scatter<<<rs_blocks, RS_WARPS * 32, 0, st>>>(P, k0, v0, pass * 8, L.nchunks, hist, k1, v1);
        uint32_t* t = k0; k0 = k1; k1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    k_box_minmax<<<L.nboxes, 256, 0, st>>>(P, points, v0, boxes);
    k_box_mean_dist<<<(P + 255) / 256, 256, 0, st>>>(P, points, v0, boxes, L.nboxes, out);
    return check_launch("gsr_dist2", false, st);
}

}  // namespace gsr

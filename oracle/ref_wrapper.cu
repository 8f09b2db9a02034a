// TEST INFRASTRUCTURE — not product code.
//
// Thin C-ABI wrapper (our code) around the UNMODIFIED reference rasterizer library
// `CudaRasterizer::Rasterizer` and `SimpleKNN::knn`, whose sources are compiled where
// they lie under /root/reference by oracle/Makefile into oracle/_ref/libref_dgr.so.
// It stands in for the reference's torch glue (DGR/rasterize_points.cu:35-230,
// KNN/spatial.cu:15-26), which only allocates tensors and forwards raw pointers; the
// wrapper does the same with grow-only cudaMalloc buffers (the moral equivalent of
// torch's caching allocator + Tensor::resize_) so the library builds in seconds with
// no torch headers.  Only tests/, bench.py's reference arm and smoke() load this.
//
// Every entry point takes raw DEVICE pointers and runs on the legacy default stream,
// exactly like the reference (rasterizer_impl.cu:290,315; forward.cu:395,436).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"
#include "simple_knn.h"

namespace {

struct GrowBuf {
    char* ptr = nullptr;
    size_t cap = 0;
    size_t size = 0;
    char* resize(size_t n) {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            size_t want = n + n / 4 + 256;
            if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; throw std::runtime_error("ref: cudaMalloc failed"); }
            cap = want;
        }
        size = n;
        return ptr;
    }
};

GrowBuf g_geom, g_bin, g_img;
int g_P = 0, g_R = 0, g_W = 0, g_H = 0;
std::string g_err;

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Mirrors RasterizeGaussiansCUDA (DGR/rasterize_points.cu:35-119): zero-fills the
// outputs, then calls Rasterizer::forward.  Null pointer == "absent" input.
// Returns num_rendered, or -1 on error.
int ref_forward(int P, int D, int M, const float* bg, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* campos,
                float tanfovx, float tanfovy, int prefiltered,
                float* out_color, float* out_depth, float* out_alpha, int* radii, int debug) {
    try {
        cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * (size_t)W * H, 0);
        cudaMemsetAsync(out_depth, 0, sizeof(float) * (size_t)W * H, 0);
        cudaMemsetAsync(out_alpha, 0, sizeof(float) * (size_t)W * H, 0);
        if (radii) cudaMemsetAsync(radii, 0, sizeof(int) * (size_t)P, 0);
        g_P = P; g_W = W; g_H = H; g_R = 0;
        if (P == 0) return 0;
        std::function<char*(size_t)> gf = [](size_t n) { return g_geom.resize(n); };
        std::function<char*(size_t)> bf = [](size_t n) { return g_bin.resize(n); };
        std::function<char*(size_t)> imf = [](size_t n) { return g_img.resize(n); };
        g_R = CudaRasterizer::Rasterizer::forward(gf, bf, imf, P, D, M, bg, W, H, means3D, shs,
                                                  colors_precomp, opacities, scales, scale_modifier,
                                                  rotations, cov3D_precomp, viewmatrix, projmatrix,
                                                  campos, tanfovx, tanfovy, prefiltered != 0,
                                                  out_color, out_depth, out_alpha, radii, debug != 0);
        return g_R;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// Mirrors RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:121-209).  The ten
// gradient buffers are caller-allocated and are zero-filled here like torch::zeros.
// Uses the geometry/binning/image buffers left by the last ref_forward.
int ref_backward(int P, int D, int M, const float* bg, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* campos, float tanfovx, float tanfovy, const int* radii,
                 const float* out_alpha, const float* dL_dcolor_img, const float* dL_ddepth_img,
                 const float* dL_dalpha_img,
                 float* dL_dmeans2D /*P*3*/, float* dL_dconic /*P*4*/, float* dL_dopacity /*P*/,
                 float* dL_dcolors /*P*3*/, float* dL_ddepths /*P*/, float* dL_dmeans3D /*P*3*/,
                 float* dL_dcov3D /*P*6*/, float* dL_dsh /*P*M*3*/, float* dL_dscales /*P*3*/,
                 float* dL_drotations /*P*4*/, int debug) {
    try {
        size_t p = (size_t)P;
        cudaMemsetAsync(dL_dmeans2D, 0, 4 * 3 * p, 0);
        cudaMemsetAsync(dL_dconic, 0, 4 * 4 * p, 0);
        cudaMemsetAsync(dL_dopacity, 0, 4 * p, 0);
        cudaMemsetAsync(dL_dcolors, 0, 4 * 3 * p, 0);
        cudaMemsetAsync(dL_ddepths, 0, 4 * p, 0);
        cudaMemsetAsync(dL_dmeans3D, 0, 4 * 3 * p, 0);
        cudaMemsetAsync(dL_dcov3D, 0, 4 * 6 * p, 0);
        if (M > 0 && dL_dsh) cudaMemsetAsync(dL_dsh, 0, 4 * 3 * p * M, 0);
        cudaMemsetAsync(dL_dscales, 0, 4 * 3 * p, 0);
        cudaMemsetAsync(dL_drotations, 0, 4 * 4 * p, 0);
        if (P == 0) return 0;
        if (P != g_P || W != g_W || H != g_H) { g_err = "ref_backward: no matching forward"; return -1; }
        CudaRasterizer::Rasterizer::backward(P, D, M, g_R, bg, W, H, means3D, shs, colors_precomp,
                                             scales, scale_modifier, rotations, cov3D_precomp,
                                             viewmatrix, projmatrix, campos, tanfovx, tanfovy, radii,
                                             g_geom.ptr, g_bin.ptr, g_img.ptr, out_alpha,
                                             dL_dcolor_img, dL_ddepth_img, dL_dalpha_img,
                                             dL_dmeans2D, dL_dconic, dL_dopacity, dL_dcolors, dL_ddepths,
                                             dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations,
                                             debug != 0);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// Device pointers into the reference's opaque workspaces after the last ref_forward, so
// parity tests can compare per-stage buffers.  Layout is re-derived with the
// reference's own fromChunk (rasterizer_impl.cu:155-193).
struct RefState {
    const float* depths; const unsigned char* clamped; const int* internal_radii;
    const float* means2D; const float* cov3D; const float* conic_opacity; const float* rgb;
    const uint32_t* point_offsets; const uint32_t* tiles_touched;
    const uint32_t* point_list; const uint64_t* point_list_keys;
    const uint32_t* ranges; const uint32_t* n_contrib;
    int P, R, W, H;
};

int ref_get_state(RefState* s) {
    memset(s, 0, sizeof(*s));
    s->P = g_P; s->R = g_R; s->W = g_W; s->H = g_H;
    if (g_P == 0) return 0;
    char* c = g_geom.ptr;
    auto geo = CudaRasterizer::GeometryState::fromChunk(c, (size_t)g_P);
    s->depths = geo.depths; s->clamped = (const unsigned char*)geo.clamped;
    s->internal_radii = geo.internal_radii; s->means2D = (const float*)geo.means2D;
    s->cov3D = geo.cov3D; s->conic_opacity = (const float*)geo.conic_opacity; s->rgb = geo.rgb;
    s->point_offsets = geo.point_offsets; s->tiles_touched = geo.tiles_touched;
    char* b = g_bin.ptr;
    auto bin = CudaRasterizer::BinningState::fromChunk(b, (size_t)g_R);
    s->point_list = bin.point_list; s->point_list_keys = bin.point_list_keys;
    char* i = g_img.ptr;
    auto img = CudaRasterizer::ImageState::fromChunk(i, (size_t)g_W * g_H);
    s->ranges = (const uint32_t*)img.ranges; s->n_contrib = img.n_contrib;
    return 0;
}

// Device-to-device copy helper so Python can snapshot the buffers ref_get_state points at.
int ref_copy(void* dst, const void* src, size_t bytes) {
    return (int)cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice);
}

// Mirrors markVisible (DGR/rasterize_points.cu:211-230).
int ref_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present) {
    cudaMemsetAsync(present, 0, (size_t)P, 0);
    if (P == 0) return 0;
    CudaRasterizer::Rasterizer::markVisible(P, (float*)means3D, (float*)viewmatrix, (float*)projmatrix,
                                            (bool*)present);
    return 0;
}

// Mirrors distCUDA2 (KNN/spatial.cu:15-26).
int ref_dist2(int P, const float* points, float* mean_dists) {
    try {
        cudaMemsetAsync(mean_dists, 0, sizeof(float) * (size_t)P, 0);
        if (P == 0) return 0;
        SimpleKNN::knn(P, (float3*)points, mean_dists);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"

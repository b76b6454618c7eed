// Shared host/device definitions for libmn_b200.so.  The whole library is compiled with
// -fmad=false: every multiply-add that the fp32 oracle performs as two separately rounded torch ops
// stays two roundings here; fused multiply-adds appear only where written explicitly as fmaf().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/mn_b200.h"

// device status bits (mn_ctx::status_d)
#define MN_STATUS_SPHERE 1u
#define MN_STATUS_OVERFLOW 2u
#define MN_STATUS_INDEX 4u

#include <vector>

// One element-wise re-layout of a weight tensor (fp32 transposes / sub-matrices / copies, fp16 tensor-core images).  The ~75 of
// them a sub-module needs are queued by mn_model_set_weights and run as TWO launches (fp32 layouts first, the fp16 images that read
// them second) instead of one tiny launch each: a training step re-packs every sub-module after the optimiser step.
enum { PK_COPY = 0, PK_TRANSPOSE, PK_SUBMATRIX, PK_TC_IMAGE, PK_TC_HALF, PK_TC_F32, PK_DGRAD, PK_RGBW };
struct PackOp {
    const float* src;
    void* dst;
    void* dst2;
    long long count;
    int kind;
    int p[7];
};

struct mn_ctx {
    int device = 0;
    int sm_count = 148;
    std::string err;
    unsigned int* status_d = nullptr;
    long long launches = 0;             // kernels launched through this context (bench: gpu_launches)
    // optional CUDA-event timing of the MLP kernel launches (bench: roofline.achieved)
    int prof_on = 0;
    std::vector<cudaEvent_t> prof_ev;   // start/stop pairs
    size_t prof_used = 0;
    // queued weight re-layouts (see PackOp) and their device-side table
    std::vector<PackOp> pack_ops;
    PackOp* pack_ops_d = nullptr;
    size_t pack_ops_cap = 0;
};
void mn_pack_push(mn_ctx* ctx, const PackOp& op);
int mn_pack_flush(mn_ctx* ctx, cudaStream_t st);

static inline void mn_prof_begin(mn_ctx* ctx, cudaStream_t st) {
    if (!ctx->prof_on) return;
    if (ctx->prof_used + 2 > ctx->prof_ev.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ctx->prof_ev.push_back(a);
        ctx->prof_ev.push_back(b);
    }
    cudaEventRecord(ctx->prof_ev[ctx->prof_used], st);
}
static inline void mn_prof_end(mn_ctx* ctx, cudaStream_t st) {
    if (!ctx->prof_on) return;
    cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], st);
    ctx->prof_used += 2;
}

static inline int mn_fail(mn_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define MN_CUDA(ctx, expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            return mn_fail(ctx, MN_ERR_CUDA,                                                        \
                           std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ +   \
                               ":" + std::to_string(__LINE__) + ")");                               \
        }                                                                                           \
    } while (0)

#define MN_LAUNCH_CHECK(ctx)               \
    do {                                   \
        (ctx)->launches++;                 \
        MN_CUDA(ctx, cudaGetLastError());  \
    } while (0)

static inline int64_t mn_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t mn_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// Where per-row network inputs come from (device-side view of mn_rows).
// ------------------------------------------------------------------------------------------------
struct RowSrc {
    const float* x;      // mode 0: [B, cols];  mode 1: xyz [B, cols]
    int cols;            // row stride of x
    int net_off;         // first network column inside x (3 when a real-xyz routing prefix is present)
    const float* dirs;   // per-ray (mode 1) or inside x (mode 0)
    int64_t dir_stride;
    const float* idx;
    int64_t idx_stride;
    int div;             // row -> ray divisor (1 in mode 0)
    int dir_quirk;       // models/nerf.py:146 x[:, -4:-1] with no index column: dir := (xyz_z, d_x, d_y)
    int xyz_dim;

    __device__ __forceinline__ float xyz(int64_t row, int j) const { return x[row * cols + net_off + j]; }
    __device__ __forceinline__ float route_xyz(int64_t row, int j) const { return x[row * cols + j]; }
    __device__ __forceinline__ float dir(int64_t row, int j) const {
        if (dir_quirk) {
            // columns [-4:-1] of [xyz(3), dir(3)] are (z, d_x, d_y)
            if (j == 0) return x[row * cols + net_off + 2];
            return dirs[(row / div) * dir_stride + (j - 1)];
        }
        return dirs[(row / div) * dir_stride + j];
    }
    __device__ __forceinline__ float index(int64_t row) const { return idx[(row / div) * idx_stride]; }
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mn_pow2f(int k) { return __int_as_float((127 + k) << 23); }

// sin / cos of (2^k * x) exactly as the oracle evaluates them: the product is exact (power of two),
// then a full-range accurate sincos (no fast-math intrinsics).
__device__ __forceinline__ void mn_pe_sincos(float x, int k, float* s, float* c) {
    sincosf(x * mn_pow2f(k), s, c);
}

// Same value, ~4x cheaper when many bands share one coordinate: reduce the argument ONCE in fp64
// (xp = x/pi, error ~1e-16 relative), then every band is an exact power-of-two scaling, an exact
// quadrant split (u = n/2 + w, |w| <= 1/4) and a short sincospi polynomial on w.  The angle error of
// rounding w to fp32 is <= pi * 2^-27 = 2.4e-8 rad, i.e. below the 1-ulp accuracy of the oracle's own sin/cos.
__device__ __forceinline__ double mn_pe_prescale(float x) { return (double)x * 0.31830988618379067154; }
__device__ __forceinline__ void mn_pe_sincos_pi(double xp, int k, float* s, float* c) {
    const double u = scalbn(xp, k);          // 2^k x / pi, exact scaling
    const double n = rint(2.0 * u);
    const float w = (float)(u - 0.5 * n);    // exact difference, then one rounding
    float sw, cw;
    sincospif(w, &sw, &cw);
    const int q = (int)((long long)n & 3);
    const bool swap = q & 1;
    const float ss = swap ? cw : sw, cc = swap ? sw : cw;
    *s = (q & 2) ? -ss : ss;                 // q: 0 (s,c) 1 (c,-s) 2 (-s,-c) 3 (-c,s)
    *c = ((q + 1) & 2) ? -cc : cc;
}

__device__ __forceinline__ float mn_softplus_shifted(float x) {
    // F.softplus(x - 1, beta=1, threshold=20)   (models/nerf.py:38-39)
    float y = x - 1.0f;
    return y > 20.0f ? y : log1pf(expf(y));
}

__device__ __forceinline__ float mn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ double mn_shfl_up_d(double v, int delta) {
    return __shfl_up_sync(0xffffffffu, v, delta);
}

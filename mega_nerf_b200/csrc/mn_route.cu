// Spatial routing of sample rows to sub-modules (models/mega_nerf.py:21-49), bucketed into
// MN_TILE-aligned slot ranges so that every MLP tile belongs to exactly one sub-module, plus the
// ascending-sub-module blend that replaces `results[mask] += sub_result * weights`.
//
// Distances restate torch.cdist's matmul path bit for bit:
//   x_ = [-2x, |x|^2, 1],  c_ = [c, 1, |c|^2],  d^2 = FMA chain over k ascending from an accumulator
//   of 0, clamp_min(0), sqrt.  Batches with <= 25 rows AND <= 25 centroids take cdist's direct path
//   sqrt(fma-accumulated sum (x-c)^2) instead, as torch does.
#include "mn_model.cuh"

namespace {

struct RoutePoint {
    float x[3];
    float xn;      // |x|^2 over the clustered dims (matmul path)
};

__device__ __forceinline__ RoutePoint load_point(const RowSrc& src, int64_t row, int s) {
    RoutePoint p;
    p.x[0] = p.x[1] = p.x[2] = 0.0f;
    for (int j = s; j < 3; ++j) p.x[j] = src.route_xyz(row, j);
    p.xn = p.x[s] * p.x[s];
    for (int j = s + 1; j < 3; ++j) p.xn = p.xn + p.x[j] * p.x[j];
    return p;
}

// distance to centroid k (cent = [K,3] in shared memory), restating torch.cdist (see file header)
__device__ __forceinline__ float dist_to(const RoutePoint& p, const float* __restrict__ cent, int k, int s, bool direct) {
    if (direct) {
        float acc = 0.0f;
        for (int j = s; j < 3; ++j) {
            const float t = p.x[j] - cent[k * 3 + j];
            acc = fmaf(t, t, acc);  // torch's direct cdist kernel contracts this (probe: 100% bitwise)
        }
        return sqrtf(acc);
    }
    float cn = cent[k * 3 + s] * cent[k * 3 + s];
    for (int j = s + 1; j < 3; ++j) cn = cn + cent[k * 3 + j] * cent[k * 3 + j];
    float acc = 0.0f;
    for (int j = s; j < 3; ++j) acc = fmaf(-2.0f * p.x[j], cent[k * 3 + j], acc);
    acc = fmaf(p.xn, 1.0f, acc);
    acc = fmaf(1.0f, cn, acc);
    return sqrtf(fmaxf(acc, 0.0f));
}

// Routing decision of one point without per-thread arrays (distances are recomputed per pass; K is small).
struct RouteDecision {
    float dmin, sum;   // minimum distance; sum of masked inverse distances (margin > 1)
    int amin;
};

__device__ __forceinline__ RouteDecision route_decide(const RoutePoint& p, const float* __restrict__ cent, int K, int s,
                                                      bool direct, float margin) {
    RouteDecision r;
    r.dmin = dist_to(p, cent, 0, s, direct);
    r.amin = 0;
    for (int k = 1; k < K; ++k) {
        const float d = dist_to(p, cent, k, s, direct);
        if (d < r.dmin) { r.dmin = d; r.amin = k; }
    }
    r.sum = 0.0f;
    if (margin > 1.0f) {
        const float thr = margin * r.dmin;
        for (int k = 0; k < K; ++k) {
            const float d = dist_to(p, cent, k, s, direct);
            float inv = 1.0f / (d + 1e-8f);
            if (d > thr) inv = 0.0f;
            r.sum = r.sum + inv;
        }
    }
    return r;
}

// blend weight of sub-module k (0 when masked out); for hard routing 1 for the argmin
__device__ __forceinline__ float route_weight(const RoutePoint& p, const RouteDecision& r, const float* __restrict__ cent, int k,
                                              int s, bool direct, float margin) {
    if (!(margin > 1.0f)) return k == r.amin ? 1.0f : 0.0f;
    const float d = dist_to(p, cent, k, s, direct);
    float inv = 1.0f / (d + 1e-8f);
    if (d > margin * r.dmin) inv = 0.0f;
    return inv / r.sum;
}

__global__ void route_count_kernel(RowSrc src, int64_t B, const float* __restrict__ cent, int K, int s, float margin,
                                   int direct, int* counters) {
    __shared__ int hist[MN_MAX_SUB];
    __shared__ float sc[MN_MAX_SUB * 3];
    for (int i = threadIdx.x; i < K * 3; i += blockDim.x) sc[i] = cent[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = row < B;
    RoutePoint p{};
    RouteDecision r{};
    if (live) {
        p = load_point(src, row, s);
        r = route_decide(p, sc, K, s, direct, margin);
    }
    for (int k = 0; k < K; ++k) {
        const bool on = live && route_weight(p, r, sc, k, s, direct, margin) > 0.0f;
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&hist[k], __popc(b));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        if (hist[i]) atomicAdd(&counters[CNT_COUNT + i], hist[i]);
}

__global__ void route_scan_kernel(int* counters, int K) {
    if (threadIdx.x == 0) {
        int off = 0, pairs = 0;
        for (int k = 0; k < K; ++k) {
            counters[CNT_START + k] = off;
            const int c = counters[CNT_COUNT + k];
            pairs += c;
            off += (c + MN_TILE - 1) / MN_TILE * MN_TILE;
            counters[CNT_CURSOR + k] = 0;
        }
        counters[CNT_START + K] = off;
        counters[CNT_NSLOTS] = off;
        counters[CNT_NPAIRS] = pairs;
    }
}

__global__ void route_scatter_kernel(RowSrc src, int64_t B, const float* __restrict__ cent, int K, int s, float margin,
                                     int direct, int* counters, int64_t cap, int* slot_row, float* slot_w,
                                     int* row_slots, unsigned int* status) {
    __shared__ float sc[MN_MAX_SUB * 3];
    for (int i = threadIdx.x; i < K * 3; i += blockDim.x) sc[i] = cent[i];
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = row < B;
    RoutePoint p{};
    RouteDecision r{};
    if (live) {
        p = load_point(src, row, s);
        r = route_decide(p, sc, K, s, direct, margin);
    }
    for (int k = 0; k < K; ++k) {
        const float w = live ? route_weight(p, r, sc, k, s, direct, margin) : 0.0f;
        const bool on = w > 0.0f;
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (!b) {
            if (live && row_slots) row_slots[row * K + k] = -1;
            continue;
        }
        int base = 0;
        const int leader = __ffs(b) - 1;
        if (lane == leader) base = atomicAdd(&counters[CNT_CURSOR + k], __popc(b));
        base = __shfl_sync(0xffffffffu, base, leader);
        int slot = -1;
        if (on) {
            slot = counters[CNT_START + k] + base + __popc(b & ((1u << lane) - 1));
            if (slot >= cap) {
                atomicOr(status, MN_STATUS_OVERFLOW);
                slot = -1;
            } else {
                slot_row[slot] = (int)row;
                if (slot_w) slot_w[slot] = w;
            }
        }
        if (live && row_slots) row_slots[row * K + k] = slot;
    }
}

__global__ void combine_kernel(int64_t B, int K, const int* __restrict__ row_slots, const float* __restrict__ slot_out,
                               int out_cols, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * out_cols) return;
    const int64_t row = i / out_cols;
    const int c = (int)(i % out_cols);
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) {  // ascending sub-module order (mega_nerf.py:34,49)
        const int slot = row_slots[row * K + k];
        if (slot >= 0) acc = acc + slot_out[(int64_t)slot * out_cols + c];
    }
    out[i] = acc;
}

__global__ void route_only_kernel(RowSrc src, int64_t B, const float* __restrict__ cent, int K, int s, float margin,
                                  int direct, int* assign, float* weights) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B) return;
    const RoutePoint p = load_point(src, row, s);
    const RouteDecision r = route_decide(p, cent, K, s, direct, margin);
    if (margin > 1.0f) {
        for (int k = 0; k < K; ++k) weights[row * K + k] = route_weight(p, r, cent, k, s, direct, margin);
    } else {
        assign[row] = r.amin;
    }
}

}  // namespace

int mn_route_build(mn_ctx* ctx, mn_model* m, const RowSrc& src, int64_t B, int64_t cap, int* slot_row, float* slot_w,
                   int* row_slots, cudaStream_t st) {
    const int K = m->d.n_sub;
    const int direct = (B <= 25 && K <= 25) ? 1 : 0;
    MN_CUDA(ctx, cudaMemsetAsync(m->counters_d, 0, CNT_TOTAL * sizeof(int), st));
    MN_CUDA(ctx, cudaMemsetAsync(slot_row, 0xFF, (size_t)cap * sizeof(int), st));
    const unsigned blocks = (unsigned)mn_cdiv(B, 256);
    route_count_kernel<<<blocks, 256, 0, st>>>(src, B, m->centroids_d, K, m->d.cluster_dim_start, m->d.boundary_margin,
                                               direct, m->counters_d);
    MN_LAUNCH_CHECK(ctx);
    route_scan_kernel<<<1, 32, 0, st>>>(m->counters_d, K);
    MN_LAUNCH_CHECK(ctx);
    route_scatter_kernel<<<blocks, 256, 0, st>>>(src, B, m->centroids_d, K, m->d.cluster_dim_start,
                                                 m->d.boundary_margin, direct, m->counters_d, cap, slot_row, slot_w,
                                                 row_slots, ctx->status_d);
    MN_LAUNCH_CHECK(ctx);
    return MN_OK;
}

int mn_route_combine(mn_ctx* ctx, mn_model* m, int64_t B, const int* row_slots, const float* slot_out, int out_cols,
                     float* out, cudaStream_t st) {
    const int64_t n = B * out_cols;
    combine_kernel<<<(unsigned)mn_cdiv(n, 256), 256, 0, st>>>(B, m->d.n_sub, row_slots, slot_out, out_cols, out);
    MN_LAUNCH_CHECK(ctx);
    return MN_OK;
}

extern "C" int mn_model_route(mn_ctx* ctx, mn_model* m, const mn_rows* rows, int64_t B, int32_t* assign_out_d,
                              float* weights_out_d, void* stream) {
    if (!ctx || !m || !rows) return MN_ERR_INVALID;
    if (m->d.kind != 2) return mn_fail(ctx, MN_ERR_INVALID, "mn_model_route: not a MegaNeRF model");
    RowSrc src{};
    src.x = rows->x_d;
    src.cols = rows->cols;
    src.div = 1;
    const int K = m->d.n_sub;
    const int direct = (B <= 25 && K <= 25) ? 1 : 0;
    if (B == 0) return MN_OK;
    route_only_kernel<<<(unsigned)mn_cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(
        src, B, m->centroids_d, K, m->d.cluster_dim_start, m->d.boundary_margin, direct, assign_out_d, weights_out_d);
    MN_LAUNCH_CHECK(ctx);
    return MN_OK;
}

// Tensor-core training path (precision 'tc_f16' for the recording forward and the backward pass), included inside
// mn_mlp_tc.cu's anonymous namespace.  SURVEY.md §8f-1; the reference trains exactly this way on a GPU: Linear layers in
// fp16 with fp32 accumulation under autocast, gradients scaled into fp16 range (runner.py:243-274, opts.py:99).
//
//   forward   tc_mlp_pp_kernel<PP_TRAIN_FWD>  the inference kernel + every layer's fp16 activations written to a tape in the
//                                             tile-image layout of the activation buffer ([cols/8][128 slots][8]).
//   dgrad     tc_mlp_pp_kernel<PP_DGRAD>      the same GEMM pipeline on transposed weight images: head stage on CUDA cores
//                                             (sigmoid' / softplus' / rgb Linear transposed), then dH_{l-1} = dZ_l W_l with the
//                                             ReLU mask read from the activation tape in the epilogue; dZ images (fp16, scaled by
//                                             a power of two S) go to a gradient tape.
//   wgrad     tc_wgrad_kernel                 dW_l = dZ_l^T X_l over all slots of a sub-module: both tapes are consumed AS THEY
//                                             ARE through MN-major UMMA descriptors (the slot axis is the K axis; LBO = 128 B
//                                             between 8-slot groups, SBO = 2048 B between 8-column groups - verified by
//                                             scripts/probes/mn_major_probe.cu), M = 128 output channels per CTA, N <= 256 input
//                                             channels + a 16-column all-ones operand whose product is the bias gradient,
//                                             accumulated in TMEM over a chunk of tiles and flushed with fp32 atomics.
//   heads     tc_heads_wgrad_kernel (sigma / rgb Linears: 1 and 3 output channels - CUDA cores), tc_emb_grad_kernel
//             (appearance embedding: W_e^T times the per-image sums of dZ_dira rows collected by the dgrad head stage).
#pragma once

// ---- data-gradient plan: GEMM chain of the backward pass as a TcPlan (all operands from the activation buffer)
inline bool build_dgrad_plan(const NetDims& nd, TcPlan* p) {
    if (nd.L != 256 || !nd.has_dir_a || nd.rgb_dim != 3 || nd.affine || nd.layers < 2 || nd.layers > 10) return false;
    TcPlan& P = *p;
    P = TcPlan{};
    P.L = nd.L;
    P.bstride = 256;
    int woff = 0, ng = 0;
    auto add = [&](int n, int k, int img, int epi) {
        TcGemm& g = P.g[ng++];
        g.n = n; g.nseg = 1; g.src[0] = SRC_H; g.k[0] = k; g.src[1] = 0; g.k[1] = 0;
        g.w_off = woff; g.bias_off = img; g.epi = epi;
        woff += k * n * 2;
    };
    add(nd.L, nd.L / 2, nd.layers, EPI_D_LINEAR);                    // dF  = dZ_dira  W_dira[:, 0:L]      -> image 'final'
    add(nd.L, nd.L, nd.layers - 1, EPI_D_MASK_SIGMA);                // dH  = dF W_final + dsigma w_sigma  -> dZ of the last trunk layer
    for (int l = nd.layers - 1; l >= 1; --l) add(nd.L, nd.L, l - 1, EPI_D_MASK);   // dZ_{l-1} = mask(dZ_l W_l[:, hidden part])
    P.n_gemm = P.n_trunk = ng;
    P.plane_bytes = woff;
    P.sigma_w_off = 0;
    P.f32_floats = nd.L + 3 * (nd.L / 2);                            // [sigma_w (L)][rgb_w [3][L/2]]
    P.f32_off = woff;
    P.sub_bytes = (int)mn_align((size_t)woff + (size_t)P.f32_floats * 4, 256);
    return true;
}

// S = 2^(10 - ceil(log2(max |g|)))  (1 if g == 0 or not finite): the gradient images hold S * dZ in fp16
__global__ void tc_grad_scale_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ scale) {
    __shared__ float red[32];
    float m = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(g[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
        float s = 1.0f;
        if (m > 0.0f && m < 3.0e38f) s = exp2f(10.0f - ceilf(log2f(m)));
        scale[0] = fminf(fmaxf(s, 1.0f / 1099511627776.0f), 1099511627776.0f);       // 2^-40 .. 2^40
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradients
// ------------------------------------------------------------------------------------------------
struct WgItem {
    int dz_off;       // byte offset of the 128-column dZ half inside a tile's gradient record
    int x_region;     // 0: activation record, 1: encoder (feature) tile
    int x_off;        // byte offset of the first X column group inside that record
    int n;            // MMA N = X columns (multiple of 16, <= 256)
    int n_real;       // columns that exist in the weight matrix
    int w_off;        // float offset of W[out0][in0] inside one sub-module's gradient block
    int k_in;         // row stride (in_features) of that weight matrix
    int b_off;        // float offset of bias[out0], or -1 when another item of the same layer owns the bias
};
constexpr int kWgMaxItems = 48;
struct WgArgs {
    WgItem item[kWgMaxItems];
    int n_items;
    const unsigned char* act;       // activation records
    const unsigned char* dz;        // gradient records (same layout)
    const unsigned char* xreg;      // encoder tiles
    int64_t act_tile_bytes, x_tile_bytes;
    const int* counters;            // routing counters saved by the forward pass, or NULL
    int64_t n_tiles;                // counters == NULL: all tiles belong to fixed_sub
    int fixed_sub;
    int chunk_tiles;
    float* gw;                      // [n_sub][sub_stride] fp32
    int64_t sub_stride;
    const float* scale;
};
constexpr int kWgStageBytes = 96 * 1024;
constexpr int kWgThreads = 192;

__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_kernel(const WgArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;                                   // 2 x 96 KiB: [dZ half 32 KiB][X <= 64 KiB]
    unsigned char* ones = smem + 2 * kWgStageBytes;               // 6 KiB of fp16 1.0
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 6144);
    uint64_t* full = bars;        // [2]
    uint64_t* empty = bars + 2;   // [2]
    uint64_t* done = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WgItem it = A.item[blockIdx.y];
    int sub = A.fixed_sub;
    int64_t t_lo = 0, t_hi = A.n_tiles;
    if (A.counters) {
        sub = (int)blockIdx.z;
        t_lo = A.counters[CNT_START + sub] / kTileM;
        t_hi = A.counters[CNT_START + sub + 1] / kTileM;
    }
    const int64_t t_begin = t_lo + (int64_t)blockIdx.x * A.chunk_tiles;
    const int64_t t_end = min(t_hi, t_begin + (int64_t)A.chunk_tiles);
    if (t_begin >= t_end) return;

    for (int i = threadIdx.x; i < 6144 / 4; i += kWgThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t xbytes = (uint32_t)(it.n / 8) * (kTileM * 16);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            const unsigned char* xbase = it.x_region ? A.xreg : A.act;
            const int64_t xstride = it.x_region ? A.x_tile_bytes : A.act_tile_bytes;
            for (int64_t t = t_begin; t < t_end; ++t) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], 32768u + xbytes);
                bulk_g2s(ring + (size_t)st * kWgStageBytes, A.dz + (size_t)t * A.act_tile_bytes + it.dz_off, 32768u, &full[st]);
                bulk_g2s(ring + (size_t)st * kWgStageBytes + 32768, xbase + (size_t)t * xstride + it.x_off, xbytes, &full[st]);
                if (++st == 2) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // A = dZ^T (M = 128 channels x K = 16 slots), B = X^T: both MN-major views of the row-major tile images
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(it.n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        const uint32_t idesc1 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        const uint64_t od = make_desc(smem_u32(ones), 128, 2048);
        uint32_t st = 0, ph = 0, accum = 0;
        for (int64_t t = t_begin; t < t_end; ++t) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t base = smem_u32(ring) + st * (uint32_t)kWgStageBytes;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t ad = make_desc(base + (uint32_t)ks * 256u, 128, 2048);
                    const uint64_t bd = make_desc(base + 32768u + (uint32_t)ks * 256u, 128, 2048);
                    tc_mma_f16(tmem_base, ad, bd, idesc, accum);
                    if (it.b_off >= 0) tc_mma_f16(tmem_base + 256u, ad, od, idesc1, accum);
                    accum = 1;
                }
                tc_commit(&empty[st]);
            }
            accum = 1;
            __syncwarp();
            if (++st == 2) { st = 0; ph ^= 1; }
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else {
        // ---- flush: lane = output channel, columns = input channels; unscale and accumulate into the fp32 gradient block
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3;                                   // TMEM lane quarter of warps 2..5: 2, 3, 0, 1
        const int r = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float inv = 1.0f / *A.scale;
        float* W = A.gw + (size_t)sub * A.sub_stride;
        for (int c0 = 0; c0 < it.n; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_lane + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < it.n_real) atomicAdd(W + it.w_off + (size_t)r * it.k_in + c0 + i, __uint_as_float(v[i]) * inv);
        }
        if (it.b_off >= 0) {
            uint32_t v[16];
            tmem_ld16(t_lane + 256u, v);
            tmem_ld_wait();
            atomicAdd(W + it.b_off + r, __uint_as_float(v[0]) * inv);
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// sigma Linear (1 x L) and rgb Linear (3 x L/2) weight / bias gradients from the fp32 head gradients and the fp16 tapes.
struct HeadsArgs {
    const unsigned char* act;
    const float* gf32;              // per tile [4][128]: d sigma pre-activation, d rgb pre-activation (3)
    int64_t act_tile_bytes;
    int L, layers;
    const int* counters;
    int64_t n_tiles;
    int fixed_sub, chunk_tiles;
    float* gw;
    int64_t sub_stride;
    int sigma_w, sigma_b, rgb_w, rgb_b;     // float offsets in a sub-module's gradient block (rgb_w is [3][L/2])
};
__global__ void __launch_bounds__(256) tc_heads_wgrad_kernel(const HeadsArgs A) {
    __shared__ float G4[4][kTileM];
    int sub = A.fixed_sub;
    int64_t t_lo = 0, t_hi = A.n_tiles;
    if (A.counters) {
        sub = (int)blockIdx.y;
        t_lo = A.counters[CNT_START + sub] / kTileM;
        t_hi = A.counters[CNT_START + sub + 1] / kTileM;
    }
    const int64_t t_begin = t_lo + (int64_t)blockIdx.x * A.chunk_tiles;
    const int64_t t_end = min(t_hi, t_begin + (int64_t)A.chunk_tiles);
    if (t_begin >= t_end) return;
    const int k = threadIdx.x, L = A.L, half = L / 2;
    float ws = 0.0f, wr0 = 0.0f, wr1 = 0.0f, wr2 = 0.0f, bs = 0.0f;
    for (int64_t t = t_begin; t < t_end; ++t) {
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * kTileM; i += 256) G4[i / kTileM][i % kTileM] = A.gf32[(size_t)t * 4 * kTileM + i];
        __syncthreads();
        const unsigned char* rec = A.act + (size_t)t * A.act_tile_bytes;
        if (k < L) {
            const __half* h = reinterpret_cast<const __half*>(rec + (size_t)(A.layers - 1) * L * kTileM * 2) + (size_t)(k >> 3) * (kTileM * 8) + (k & 7);
            for (int r = 0; r < kTileM; ++r) ws = fmaf(G4[0][r], __half2float(h[r * 8]), ws);
        }
        if (k < half) {
            const __half* g = reinterpret_cast<const __half*>(rec + (size_t)(A.layers + 1) * L * kTileM * 2) + (size_t)(k >> 3) * (kTileM * 8) + (k & 7);
            for (int r = 0; r < kTileM; ++r) {
                const float gv = __half2float(g[r * 8]);
                wr0 = fmaf(G4[1][r], gv, wr0);
                wr1 = fmaf(G4[2][r], gv, wr1);
                wr2 = fmaf(G4[3][r], gv, wr2);
            }
        }
        if (k < 4) for (int r = 0; r < kTileM; ++r) bs += G4[k][r];
    }
    float* W = A.gw + (size_t)sub * A.sub_stride;
    if (k < L) atomicAdd(W + A.sigma_w + k, ws);
    if (k < half) {
        atomicAdd(W + A.rgb_w + k, wr0);
        atomicAdd(W + A.rgb_w + half + k, wr1);
        atomicAdd(W + A.rgb_w + 2 * half + k, wr2);
    }
    if (k == 0) atomicAdd(W + A.sigma_b, bs);
    else if (k < 4) atomicAdd(W + A.rgb_b + (k - 1), bs);
}

// embedding_a.weight[id][j] += sum_k We[k][j] * S[sub][id][k]     (We = dir_a_encoding columns of the embedding, [L/2][app])
__global__ void tc_emb_grad_kernel(const float* __restrict__ emb_sum, const float* __restrict__ packed_bwd, int64_t bwd_stride, int dira_e,
                                   int half, int app, int app_count, float* gw, int64_t sub_stride, int emb_off) {
    const int id = blockIdx.x, sub = blockIdx.y, j = threadIdx.x;
    if (j >= app) return;
    const float* S = emb_sum + ((size_t)sub * app_count + id) * half;
    const float* We = packed_bwd + (size_t)sub * bwd_stride + dira_e;
    float acc = 0.0f;
    bool any = false;
    for (int k = 0; k < half; ++k) {
        const float s = S[k];
        any |= s != 0.0f;
        acc = fmaf(We[k * app + j], s, acc);
    }
    if (any) atomicAdd(gw + (size_t)sub * sub_stride + emb_off + (size_t)id * app + j, acc);
}

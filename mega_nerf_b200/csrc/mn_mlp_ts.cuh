// TS variant of the tensor-core MLP (included inside mn_mlp_tc.cu's anonymous namespace).
//
// Shared-memory bandwidth, not the tensor pipe, bounds the SS kernels: per 128x256x256 layer they move
// 64 KiB (A read) + 128 KiB (B read) + 128 KiB (TMA weight fill) + 64 KiB (epilogue stores) through a
// 128 B/cycle port, i.e. 3072 cycles for 2048 cycles of MMA.  Here the activations never touch shared memory:
//   * the A operand of every hidden layer lives in TENSOR MEMORY (tcgen05.mma with A from TMEM) and is written
//     there by the epilogue with tcgen05.st (fp16 pairs, one TMEM lane per row);
//   * every layer is issued as two N=128 halves with separate 128-column accumulators, so the epilogue of half 0
//     runs under the MMAs of half 1 and the epilogue of half 1 under the next layer's half 0 (which only needs
//     the K-slabs published so far):   TMEM columns  [0,128) acc half 0 | [128,256) acc half 1 | [256,384) A0 | [384,512) A1.
// Only the weights (B) and the 80-column encodings (first layer, skip layer, view/appearance layer) use shared memory.
#pragma once

constexpr int kTsStageCols = 64;                       // K-columns per ring stage
constexpr int kTsStageBytes = kTsStageCols * 128 * 2;  // x 128 weight rows (one N-half) x fp16 = 16 KiB
constexpr int kTsMaxStages = 10;                      // split into two rings of 5 (one per N-half / issuing warp)
constexpr int kTsThreads = 640;                        // 16 epilogue warps + 2 producers + 2 MMA issuers
constexpr int kTsWarpProd0 = 16, kTsWarpProd1 = 17, kTsWarpMma0 = 18, kTsWarpMma1 = 19;

struct TsLayout {
    int ring, stages, xa, f32, sigp, bars, total;
};

__host__ __device__ inline TsLayout ts_layout(const TcPlan& p) {
    TsLayout s;
    const int kx = p.kpe > p.kaux ? p.kpe : p.kaux;
    const int f32b = ((p.f32_floats * 4 + 15) / 16) * 16;
    const int fixed = kx * kTileM * 2 + f32b + 4096 + 512;
    int st = (kSmemMax - fixed) / kTsStageBytes;
    if (st > kTsMaxStages) st = kTsMaxStages;
    s.stages = st;
    s.ring = 0;
    s.xa = st * kTsStageBytes;
    s.f32 = s.xa + kx * kTileM * 2;
    s.sigp = s.f32 + f32b;
    s.bars = s.sigp + 4096;      // [8][128] partial sigma sums
    s.total = s.bars + 512;
    return s;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Two MMA-issuing warps (one per N-half, each with its own weight ring and producer warp): a single issuing
// thread needs ~170 cycles of scalar work per tcgen05.mma and cannot keep the tensor pipe (64 cycles per N=128 MMA) fed.
__global__ void __launch_bounds__(kTsThreads, 1) tc_mlp_ts_kernel(const TcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const TsLayout SL = ts_layout(P);
    const int kStages = SL.stages / 2;            // per ring
    unsigned char* ring_all = smem + SL.ring;
    unsigned char* XA = smem + SL.xa;
    float* F32 = reinterpret_cast<float*>(smem + SL.f32);
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;                        // [10]
    uint64_t* empty = bars + kTsMaxStages;        // [10]
    uint64_t* xa_full = bars + 2 * kTsMaxStages;  // 20
    uint64_t* xa_empty = xa_full + 1;             // 21
    uint64_t* acc_full = xa_full + 2;             // [2] 22,23
    uint64_t* acc_free = xa_full + 4;             // [2] 24,25
    uint64_t* aready = xa_full + 6;               // [4] 26..29
    uint64_t* f32_full = xa_full + 10;            // 30
    uint64_t* f32_empty = xa_full + 11;           // 31
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xa_full + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTsMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(xa_full, 1);
        mbar_init(xa_empty, 2);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], kEpiWarps); }
        for (int i = 0; i < 4; ++i) mbar_init(&aready[i], 8);
        mbar_init(f32_full, 1);
        mbar_init(f32_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTsWarpProd0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto sub_of = [&](int64_t tile) -> int {
        int sub = A.m.fixed_sub;
        if (A.m.counters) {
            sub = 0;
            const int64_t s0 = tile * kTileM;
            while (sub + 1 < A.m.n_sub && s0 >= A.m.counters[CNT_START + sub + 1]) ++sub;
        }
        return sub;
    };
    // byte offset of the half-major TS weight plane inside one sub-module's pack
    const size_t ts_off = (size_t)P.plane_bytes * 2 + (size_t)(((P.f32_floats * 4 + 255) / 256) * 256);

    if (warp == kTsWarpProd0 || warp == kTsWarpProd1) {
        // =========================== TMA producers (ring `me` feeds N-half `me`) ===========================
        const int me = warp == kTsWarpProd1 ? 1 : 0;
        unsigned char* ring = ring_all + (size_t)me * kStages * kTsStageBytes;
        uint64_t* full_me = full + me * (kTsMaxStages / 2);
        uint64_t* empty_me = empty + me * (kTsMaxStages / 2);
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, xphase = 0, fphase = 0;
            const uint32_t f32_bytes = (uint32_t)(((P.f32_floats * 4 + 15) / 16) * 16);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const unsigned char* wsub = A.wpack + (size_t)sub_of(tile) * P.sub_bytes;
                if (me == 0) {
                    mbar_wait(f32_empty, fphase ^ 1);
                    mbar_expect_tx(f32_full, f32_bytes);
                    bulk_g2s(F32, wsub + (size_t)P.plane_bytes * 2, f32_bytes, f32_full);
                    fphase ^= 1;
                }
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int nw = g.n < 128 ? g.n : 128;
                    const int nh = (g.n + 127) / 128;
                    const int K = g.k[0] + (g.nseg > 1 ? g.k[1] : 0);
                    for (int h = 0; h < nh; ++h) {
                        if (h != me) continue;
                        const unsigned char* wimg = wsub + ts_off + g.w_off + (size_t)h * K * nw * 2;
                        int kbase = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            if (g.src[sgi] != SRC_H && me == 0) {
                                const __half* xt = A.ximg + tile * (int64_t)(P.kpe + P.kaux) * kTileM +
                                                   (g.src[sgi] == SRC_XAUX ? (int64_t)P.kpe * kTileM : 0);
                                mbar_wait(xa_empty, xphase ^ 1);
                                mbar_expect_tx(xa_full, (uint32_t)(kseg * kTileM * 2));
                                bulk_g2s(XA, xt, (uint32_t)(kseg * kTileM * 2), xa_full);
                                xphase ^= 1;
                            }
                            for (int k0 = 0; k0 < kseg; k0 += kTsStageCols) {
                                const int kc = min(kTsStageCols, kseg - k0);
                                const uint32_t bytes = (uint32_t)(kc * nw * 2);
                                mbar_wait(&empty_me[stage], phase ^ 1);
                                mbar_expect_tx(&full_me[stage], bytes);
                                bulk_g2s(ring + (size_t)stage * kTsStageBytes, wimg + (size_t)(kbase + k0) * nw * 2, bytes, &full_me[stage]);
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                            }
                            kbase += kseg;
                        }
                    }
                }
            }
        }
    } else if (warp == kTsWarpMma0 || warp == kTsWarpMma1) {
        // =========================== MMA issuers (whole warp; one elected lane issues) ===========================
        const int me = warp == kTsWarpMma1 ? 1 : 0;
        unsigned char* ring = ring_all + (size_t)me * kStages * kTsStageBytes;
        uint64_t* full_me = full + me * (kTsMaxStages / 2);
        uint64_t* empty_me = empty + me * (kTsMaxStages / 2);
        int stage = 0;
        uint32_t phase = 0, xphase = 0, gidx = 0;
        uint32_t fph0 = 0, fph1 = 0, rph0 = 0, rph1 = 0, rph2 = 0, rph3 = 0;
        bool used0 = false, used1 = false;
        const uint32_t xa_base = smem_u32(XA), ring_base = smem_u32(ring);
        const uint32_t full_a = smem_u32(full_me), empty_a = smem_u32(empty_me);
        const uint32_t xa_full_a = smem_u32(xa_full), xa_empty_a = smem_u32(xa_empty);
        const uint32_t acc_full_a = smem_u32(acc_full), acc_free_a = smem_u32(acc_free), aready_a = smem_u32(aready);
        const uint64_t a_step = (uint64_t)((2 * kTileM * 16) >> 4);
        const uint64_t st_step = (uint64_t)(kTsStageBytes >> 4);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int gi = 0; gi < n_gemm; ++gi, ++gidx) {
                const TcGemm& g = P.g[gi];
                const int nw = g.n < 128 ? g.n : 128;
                const int nh = (g.n + 127) / 128;
                const uint32_t idesc = make_idesc(nw);
                const uint64_t b_step = (uint64_t)((2 * nw * 16) >> 4);
                const uint64_t bd0 = make_desc(ring_base, (uint32_t)nw * 16, 128);
                // A operand of this GEMM (written by the previous epilogue): buffer (gidx-1)&1  ==  (gidx+1)&1
                const uint32_t a_tm = tmem_base + 256u + ((gidx + 1u) & 1u) * 128u;
                if (nh == 1 && me == 1) {
                    // nothing to issue for a single-half GEMM, but the shared encoding buffer is released by both issuers
                    for (int sgi = 0; sgi < g.nseg; ++sgi) {
                        if (g.src[sgi] != SRC_H) {
                            mbar_wait_a(xa_full_a, xphase);
                            xphase ^= 1;
                            commit_elect(xa_empty_a);
                        } else {
                            // keep this warp's parities of the A-slab barriers in step with the publications it skips
                            const int ns = (g.k[sgi] + 63) >> 6;
                            if (ns > 0) rph0 ^= 1;
                            if (ns > 1) rph1 ^= 1;
                            if (ns > 2) rph2 ^= 1;
                            if (ns > 3) rph3 ^= 1;
                        }
                    }
                    continue;
                }
                for (int h = 0; h < nh; ++h) {
                    if (h != me) continue;
                    // accumulator half h must have been drained by the previous epilogue that used it
                    if (h == 0) { if (used0) { mbar_wait_a(acc_free_a, fph0); fph0 ^= 1; } used0 = true; }
                    else        { if (used1) { mbar_wait_a(acc_free_a + 8, fph1); fph1 ^= 1; } used1 = true; }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)h * 128u;
                    uint32_t accum = 0;
                    for (int sgi = 0; sgi < g.nseg; ++sgi) {
                        const int kseg = g.k[sgi];
                        const bool from_x = g.src[sgi] != SRC_H;
                        if (from_x) {
                            mbar_wait_a(xa_full_a, xphase);
                            xphase ^= 1;
                            tc_fence_after();
                        }
                        // two ring stages (up to 128 K-columns, 8 MMAs) per iteration; all barrier probes go out together
                        for (int k0 = 0; k0 < kseg; k0 += 2 * kTsStageCols) {
                            const int kc0 = min(kTsStageCols, kseg - k0);
                            const int kc1 = min(kTsStageCols, kseg - k0 - kTsStageCols);     // <= 0: single stage
                            const bool two = kc1 > 0;
                            const int st0 = stage;
                            const uint32_t ph0 = phase;
                            int st1 = stage + 1;
                            uint32_t ph1 = phase;
                            if (st1 == kStages) { st1 = 0; ph1 ^= 1; }
                            const uint32_t f0 = full_a + 8u * (uint32_t)st0, f1 = two ? full_a + 8u * (uint32_t)st1 : f0;
                            const uint32_t fp1 = two ? ph1 : ph0;
                            if (!from_x) {
                                // K-slabs k0/64 (and the next one) of the A operand have been published by the previous epilogue
                                const int s0 = k0 >> 6;
                                const uint32_t r0 = s0 == 0 ? rph0 : rph2;      // s0 is 0 or 2
                                const uint32_t r1 = s0 == 0 ? rph1 : rph3;
                                const uint32_t ra0 = aready_a + 8u * (uint32_t)s0, ra1 = two ? ra0 + 8u : ra0;
                                mbar_wait4(ra0, r0, ra1, two ? r1 : r0, f0, ph0, f1, fp1);
                                if (s0 == 0) { rph0 ^= 1; if (two) rph1 ^= 1; } else { rph2 ^= 1; if (two) rph3 ^= 1; }
                            } else {
                                mbar_wait4(f0, ph0, f1, fp1, f0, ph0, f1, fp1);
                            }
                            tc_fence_after();
                            const uint64_t bdA = bd0 + (uint64_t)st0 * st_step, bdB = bd0 + (uint64_t)st1 * st_step;
                            if (from_x) {
                                ts_stage_smem(d_tmem, make_desc(xa_base + (uint32_t)(k0 / 8) * (kTileM * 16), kTileM * 16, 128), a_step, bdA,
                                              b_step, idesc, accum, kc0 >> 4, empty_a + 8u * (uint32_t)st0);
                                if (two)
                                    ts_stage_smem(d_tmem, make_desc(xa_base + (uint32_t)((k0 + kTsStageCols) / 8) * (kTileM * 16), kTileM * 16, 128),
                                                  a_step, bdB, b_step, idesc, 1u, kc1 >> 4, empty_a + 8u * (uint32_t)st1);
                            } else {
                                ts_stage_tmem(d_tmem, a_tm + (uint32_t)(k0 >> 1), bdA, b_step, idesc, accum, kc0 >> 4,
                                              empty_a + 8u * (uint32_t)st0);
                                if (two)
                                    ts_stage_tmem(d_tmem, a_tm + (uint32_t)((k0 + kTsStageCols) >> 1), bdB, b_step, idesc, 1u, kc1 >> 4,
                                                  empty_a + 8u * (uint32_t)st1);
                            }
                            accum = 1;
                            if (two) { stage = st1; phase = ph1; }
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                        }
                        if (from_x) commit_elect(xa_empty_a);
                    }
                    commit_elect(acc_full_a + 8u * (uint32_t)h);
                }
            }
        }
    } else {
        // =========================== epilogue (16 warps) ===========================
        const int q = warp & 3;                      // TMEM lane quarter
        const int part = warp >> 2;                  // 32-column piece of the 128-column accumulator half
        const int r = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t aph0 = 0, aph1 = 0, fphase = 0, gidx = 0;
        const int L = P.L;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t slot = tile * kTileM + r;
            int64_t row = -1;
            if (slot < n_slots) row = A.m.slot_row ? (int64_t)A.m.slot_row[slot] : slot;
            mbar_wait(f32_full, fphase);
            fphase ^= 1;
            float sigma = 0.0f, sacc = 0.0f;
            for (int gi = 0; gi < n_gemm; ++gi, ++gidx) {
                const TcGemm& g = P.g[gi];
                const int nw = g.n < 128 ? g.n : 128;
                const int nh = (g.n + 127) / 128;
                const float* bias = F32 + g.bias_off;
                const float* sw = F32 + P.sigma_w_off;
                const bool relu = g.epi != EPI_LINEAR;
                const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                const bool publish = g.epi != EPI_RGB && !(want_sigma && A.m.sigma_only);
                const uint32_t a_next = t_lane + 256u + (gidx & 1u) * 128u;   // this epilogue writes buffer gidx&1
                for (int h = 0; h < nh; ++h) {
                    if (h == 0) { mbar_wait(&acc_full[0], aph0); aph0 ^= 1; }
                    else        { mbar_wait(&acc_full[1], aph1); aph1 ^= 1; }
                    tc_fence_after();
                    const int c0 = 32 * part;               // column inside the half
                    const int n0 = 128 * h + c0;            // output channel == next layer's K index
                    const bool active = c0 < nw;
                    uint32_t v[32];
                    if (active) {
                        tmem_ld32(t_lane + (uint32_t)h * 128u + (uint32_t)c0, v);
                        tmem_ld_wait();
                    }
                    // this warp no longer needs accumulator half h
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_free[h]);
                    if (g.epi == EPI_RGB) {
                        if (part == 0 && row >= 0) {
                            const NetDims& nd = A.m.nd;
                            const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                            const float w = A.m.slot_w ? A.m.slot_w[slot] : 1.0f;
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                if (c < nd.rgb_dim) {
                                    float x = __uint_as_float(v[c]) + bias[c];
                                    if (nd.rgb_dim == 3) x = mn_sigmoid(x);
                                    A.m.out[o + c] = A.m.slot_w ? x * w : x;
                                }
                            }
                            A.m.out[o + nd.rgb_dim] = A.m.slot_w ? sigma * w : sigma;
                        }
                        continue;
                    }
                    if (active) {
                        float f[32];
                        const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b = b4[i];
                            f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b.x;
                            f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b.y;
                            f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b.z;
                            f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b.w;
                        }
                        if (relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
                        }
                        if (want_sigma) {
                            const float4* s4 = reinterpret_cast<const float4*>(sw + n0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 s = s4[i];
                                sacc = fmaf(f[4 * i + 0], s.x, sacc);
                                sacc = fmaf(f[4 * i + 1], s.y, sacc);
                                sacc = fmaf(f[4 * i + 2], s.z, sacc);
                                sacc = fmaf(f[4 * i + 3], s.w, sacc);
                            }
                        }
                        if (publish) {
                            uint32_t hp[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) hp[e] = pack_h2(f[2 * e], f[2 * e + 1]);
                            tmem_st16(a_next + (uint32_t)(n0 >> 1), hp);
                            tmem_st_wait();
                        }
                    }
                    if (publish && active) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&aready[2 * h + (part >> 1)]);
                    }
                }
                if (want_sigma) {
                    SIGP[part * kTileM + r] = sacc;
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (part == 0) {
                        float s = ((SIGP[r] + SIGP[kTileM + r]) + (SIGP[2 * kTileM + r] + SIGP[3 * kTileM + r])) + sw[L];
                        if (A.m.sigma_noise && row >= 0) s = s + A.m.sigma_noise[row];
                        sigma = A.m.nd.softplus ? mn_softplus_shifted(s) : fmaxf(s, 0.0f);
                        if (A.m.sigma_only && row >= 0) {
                            const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                            A.m.out[o] = A.m.slot_w ? sigma * A.m.slot_w[slot] : sigma;
                        }
                    }
                    sacc = 0.0f;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(f32_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTsWarpProd0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

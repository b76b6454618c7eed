// CTA-pair variant of the ping-pong MLP kernel (included inside mn_mlp_tc.cu's anonymous namespace).
//
// A cluster of two CTAs (one SM pair) runs tcgen05.mma.cta_group::2: one instruction issued by the leader CTA drives
// both SMs' tensor cores on a 256-row tile (128 rows per CTA).  Each CTA stages only HALF of every weight slab
// (N/2 rows of B), so per SM the TMA fill traffic, the B reads from shared memory and — decisive for this kernel —
// the scalar issue work per FLOP are all halved, and a 64-column stage (4 MMAs per barrier wait) fits the same 16 KiB.
// Everything else follows tc_mlp_pp_kernel: two 256-row tiles (X, Y) per cluster, GEMMs issued X_l, Y_l, X_l+1, ...,
// each CTA's 16 epilogue warps drain its own 128 TMEM lanes under the other tile's MMAs.
//
// Cross-CTA signalling (all barriers live at identical offsets in both CTAs):
//   full[s], xa_full   leader's barrier counts both producers: the peer arrives/expect_tx's and lets its TMA
//                      complete_tx on the LEADER's barrier (shared::cluster address from mapa);
//   empty[s], xa_empty, acc_full[slot]   tcgen05.commit ... multicast::cluster to both CTAs;
//   epi_done[slot]     leader's barrier, 32 arrivals (16 epilogue warps per CTA, the peer's arrive remotely).
#pragma once


constexpr int kC2Stages = 3;
constexpr int kC2StageCols = 64;
constexpr int kC2StageBytes = kC2StageCols * 128 * 2;   // 64 K-columns x 128 weight rows (half of N = 256) x fp16

struct C2Layout {
    int ring, h, xa, f32, f32_stride, sigp, bars, total;
};

__host__ __device__ inline C2Layout c2_layout(const TcPlan& p) {
    C2Layout s;
    const int kx = p.kpe > p.kaux ? p.kpe : p.kaux;
    s.ring = 0;
    s.h = kC2Stages * kC2StageBytes;
    s.xa = s.h + 2 * p.L * kTileM * 2;
    s.f32 = s.xa + kx * kTileM * 2;
    s.f32_stride = ((p.f32_floats * 4 + 15) / 16) * 16;
    s.sigp = s.f32 + 2 * s.f32_stride;
    s.bars = s.sigp + 2048;
    s.total = s.bars + 256;
    return s;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// TMA bulk copy into THIS CTA's shared memory, completion signalled on a (possibly remote) cluster barrier address
__device__ __forceinline__ void bulk_g2s_cbar(void* dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(bar_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar_addr, uint32_t parity) {
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("mn_mlp_tc: cluster mbarrier timeout block %d\n", (int)blockIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// up to four K=16 steps of a 256-row (2-CTA) MMA against one ring stage, then release the stage in both CTAs
__device__ __forceinline__ void c2_stage(uint32_t d_tmem, uint64_t ad, uint64_t a_step, uint64_t bd, uint64_t b_step, uint32_t idesc,
                                         uint32_t accum, int nk, uint32_t empty_bar) {
    asm volatile(
        "{\n\t.reg .pred e, p, q1, q2, q3;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t.reg .b16 msk;\n\t"
        "mov.b16 msk, 3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.gt.and.s32 q1, %7, 1, e;\n\t"
        "setp.gt.and.s32 q2, %7, 2, e;\n\t"
        "setp.gt.and.s32 q3, %7, 3, e;\n\t"
        "add.u64 a1, %1, %2;\n\tadd.u64 a2, a1, %2;\n\tadd.u64 a3, a2, %2;\n\t"
        "add.u64 b1, %3, %4;\n\tadd.u64 b2, b1, %4;\n\tadd.u64 b3, b2, %4;\n\t"
        "@e  tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %3, %5, p;\n\t"
        "@q1 tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %5, 1;\n\t"
        "@q2 tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %5, 1;\n\t"
        "@q3 tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %5, 1;\n\t"
        "@e  tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], msk;\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(a_step), "l"(bd), "l"(b_step), "r"(idesc), "r"(accum), "r"(nk), "r"(empty_bar)
        : "memory");
}
__device__ __forceinline__ void c2_commit_both(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t.reg .b16 msk;\n\t"
        "mov.b16 msk, 3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], msk;\n\t}"
        ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_m(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 2-D tensor TMA (rows of 256 B) into THIS CTA's shared memory; with .cta_group::2 the completion may be signalled on
// the peer CTA's mbarrier (cluster address), which the non-tensor bulk copy cannot do.
__device__ __forceinline__ void tma2d_c2(uint32_t dst_smem, const CUtensorMap* tm, int row, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"(tm), "r"(0), "r"(row), "r"(bar_cluster_addr)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    tc_mlp_c2_kernel(const TcArgs A, const __grid_constant__ CUtensorMap tm_big, const __grid_constant__ CUtensorMap tm_small) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const C2Layout SL = c2_layout(P);
    unsigned char* ring = smem + SL.ring;
    unsigned char* Hs = smem + SL.h;
    unsigned char* XA = smem + SL.xa;
    float* F32 = reinterpret_cast<float*>(smem + SL.f32);
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;            // [3]   per CTA: this CTA's weight-half has landed
    uint64_t* empty = bars + 3;       // [3]   per CTA (released by the leader's multicast commit)
    uint64_t* pfull = bars + 6;       // [3]   leader's: the PEER's weight-half has landed (relayed by the peer's warp 17)
    uint64_t* xa_full = bars + 9;     //       per CTA
    uint64_t* xa_empty = bars + 10;   //       per CTA
    uint64_t* pxa_full = bars + 11;   //       leader's: relayed
    uint64_t* acc_full = bars + 12;   // [2]   per CTA
    uint64_t* epi_done = bars + 14;   // [2]   leader's
    uint64_t* f32_full = bars + 16;   // [2]   per CTA
    uint64_t* f32_empty = bars + 18;  // [2]   per CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;        // 128-row tiles
    const int64_t n_super = (n_tiles + 1) / 2;                       // 256-row tiles (one per CTA pair)
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;
    const int h_bytes = P.L * kTileM * 2;
    const int64_t cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;        // cluster index / count
    const int64_t stride2 = 2 * ncl;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kC2Stages; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); mbar_init(&pfull[i], 1); }
        mbar_init(xa_full, 1);
        mbar_init(xa_empty, 1);
        mbar_init(pxa_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&epi_done[i], 2 * kEpiWarps);
            mbar_init(&f32_full[i], 1);
            mbar_init(&f32_empty[i], kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpProd) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
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
    // per sub-module pack: [hi][lo][f32 block][TS plane][cta-pair plane]
    const size_t c2_off = (size_t)P.plane_bytes * 3 + (size_t)(((P.f32_floats * 4 + 255) / 256) * 256);

    if (warp == kWarpProd) {
        // =========================== TMA producer (each CTA streams its N-half of every weight slab) ===========================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, xphase = 0, fph[2] = {0, 0};
            const uint32_t f32_bytes = (uint32_t)SL.f32_stride;
            const uint32_t full_l = mapa_u32(smem_u32(full), 0);          // the LEADER's full[] barriers count both weight halves
            const uint32_t ring_a = smem_u32(ring);
            for (int64_t t0 = cl; t0 < n_super; t0 += stride2) {
                const int64_t sup[2] = {t0, t0 + ncl};
                const unsigned char* wsub[2] = {nullptr, nullptr};
                for (int sl = 0; sl < 2; ++sl) {
                    if (sup[sl] >= n_super) continue;
                    wsub[sl] = A.wpack + (size_t)sub_of(2 * sup[sl]) * P.sub_bytes;
                    mbar_wait(&f32_empty[sl], fph[sl] ^ 1);
                    mbar_expect_tx(&f32_full[sl], f32_bytes);
                    bulk_g2s(reinterpret_cast<unsigned char*>(F32) + (size_t)sl * SL.f32_stride,
                             wsub[sl] + (size_t)P.plane_bytes * 2, f32_bytes, &f32_full[sl]);
                    fph[sl] ^= 1;
                }
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int nhalf = g.n >> 1;                                   // weight rows held by this CTA
                    const int K = g.k[0] + (g.nseg > 1 ? g.k[1] : 0);
                    for (int sl = 0; sl < 2; ++sl) {
                        if (!wsub[sl]) continue;
                        const unsigned char* wimg = wsub[sl] + c2_off + g.w_off + (size_t)rank * K * nhalf * 2;
                        const int64_t my_tile = 2 * sup[sl] + rank;              // this CTA's 128-row tile
                        int kbase = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            if (g.src[sgi] != SRC_H) {
                                const __half* xt = A.ximg + my_tile * (int64_t)(P.kpe + P.kaux) * kTileM +
                                                   (g.src[sgi] == SRC_XAUX ? (int64_t)P.kpe * kTileM : 0);
                                mbar_wait(xa_empty, xphase ^ 1);
                                mbar_expect_tx(xa_full, (uint32_t)(kseg * kTileM * 2));
                                bulk_g2s(XA, xt, (uint32_t)(kseg * kTileM * 2), xa_full);
                                xphase ^= 1;
                            }
                            for (int k0 = 0; k0 < kseg; k0 += kC2StageCols) {
                                const int kc = min(kC2StageCols, kseg - k0);
                                const uint32_t bytes = (uint32_t)(kc * nhalf * 2);
                                mbar_wait(&empty[stage], phase ^ 1);
                                mbar_expect_tx_cluster(full_l + 8u * (uint32_t)stage, bytes);
                                const int row0 = (int)((size_t)((wimg + (size_t)(kbase + k0) * nhalf * 2) - A.wpack) >> 8);
                                const uint32_t dst = ring_a + (uint32_t)stage * kC2StageBytes;
                                if (bytes == 16384u) {
                                    tma2d_c2(dst, &tm_big, row0, full_l + 8u * (uint32_t)stage);
                                } else {
                                    for (uint32_t b = 0; b < bytes; b += 2048u)
                                        tma2d_c2(dst + b, &tm_small, row0 + (int)(b >> 8), full_l + 8u * (uint32_t)stage);
                                }
                                if (++stage == kC2Stages) { stage = 0; phase ^= 1; }
                            }
                            kbase += kseg;
                        }
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer: leader CTA only, whole warp, one elected lane issues ===========================
        if (leader) {
            int stage = 0;
            uint32_t phase = 0, xphase = 0, eph0 = 0, eph1 = 0;
            bool started0 = false, started1 = false;
            const uint32_t h_base = smem_u32(Hs), xa_base = smem_u32(XA), ring_base = smem_u32(ring);
            const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), pfull_a = smem_u32(pfull);
            const uint32_t xa_full_a = smem_u32(xa_full), xa_empty_a = smem_u32(xa_empty), pxa_full_a = smem_u32(pxa_full);
            const uint32_t acc_full_a = smem_u32(acc_full), epi_done_a = smem_u32(epi_done);
            const uint64_t a_step = (uint64_t)((2 * kTileM * 16) >> 4);
            const uint64_t st_step = (uint64_t)(kC2StageBytes >> 4);
            for (int64_t t0 = cl; t0 < n_super; t0 += stride2) {
                const bool valid1 = t0 + ncl < n_super;
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int nhalf = g.n >> 1;
                    const uint32_t idesc = make_idesc_m(256, g.n);
                    const uint64_t b_step = (uint64_t)((2 * nhalf * 16) >> 4);
                    const uint64_t bd0 = make_desc(ring_base, (uint32_t)nhalf * 16, 128);
                    for (int sl = 0; sl < 2; ++sl) {
                        if (sl == 1 && !valid1) continue;
                        // both CTAs have drained this slot's accumulator and written its activations
                        if (sl == 0) { if (started0) { mbar_wait_cluster(epi_done_a, eph0); eph0 ^= 1; } started0 = true; }
                        else         { if (started1) { mbar_wait_cluster(epi_done_a + 8, eph1); eph1 ^= 1; } started1 = true; }
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)sl * 256u;
                        uint32_t accum = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            const bool from_x = g.src[sgi] != SRC_H;
                            uint64_t ad = make_desc(from_x ? xa_base : h_base + (uint32_t)(sl * h_bytes), kTileM * 16, 128);
                            if (from_x) {
                                mbar_wait_a(xa_full_a, xphase);
                                mbar_wait_cluster(pxa_full_a, xphase);
                                xphase ^= 1;
                            }
                            for (int k0 = 0; k0 < kseg; k0 += kC2StageCols) {
                                const int kc = min(kC2StageCols, kseg - k0);
                                mbar_wait_cluster(full_a + 8u * (uint32_t)stage, phase);
                                tc_fence_after();
                                c2_stage(d_tmem, ad, a_step, bd0 + (uint64_t)stage * st_step, b_step, idesc, accum, kc >> 4,
                                         empty_a + 8u * (uint32_t)stage);
                                accum = 1;
                                ad += (uint64_t)(kc >> 4) * a_step;
                                if (++stage == kC2Stages) { stage = 0; phase ^= 1; }
                            }
                            if (from_x) c2_commit_both(xa_empty_a);
                        }
                        c2_commit_both(acc_full_a + 8u * (uint32_t)sl);
                    }
                }
            }
        } else if (lane == 0) {
            // peer CTA: relay "my half of this stage has landed" to the leader, in the exact consumption order
            uint32_t xphase = 0;
            const uint32_t pxa_full_l = mapa_u32(smem_u32(pxa_full), 0);
            for (int64_t t0 = cl; t0 < n_super; t0 += stride2) {
                const bool valid1 = t0 + ncl < n_super;
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    for (int sl = 0; sl < 2; ++sl) {
                        if (sl == 1 && !valid1) continue;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            if (g.src[sgi] != SRC_H) {
                                mbar_wait(xa_full, xphase);
                                xphase ^= 1;
                                mbar_arrive_cluster(pxa_full_l);
                            }
                        }
                    }
                }
            }
        }
    } else {
        // =========================== epilogue (16 warps per CTA, own 128 TMEM lanes) ===========================
        const int q = warp & 3;
        const int part = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t aph0 = 0, aph1 = 0, fph0 = 0, fph1 = 0;
        const int L = P.L;
        const uint32_t epi_done_l = mapa_u32(smem_u32(epi_done), 0);      // leader's barrier (local address if we are the leader)
        for (int64_t t0 = cl; t0 < n_super; t0 += stride2) {
            const bool valid1 = t0 + ncl < n_super;
            int64_t slot_[2] = {0, 0}, row_[2] = {-1, -1};
            float sigma_[2] = {0.0f, 0.0f};
            for (int sl = 0; sl < 2; ++sl) {
                if (sl == 1 && !valid1) continue;
                slot_[sl] = (2 * (t0 + (int64_t)sl * ncl) + rank) * kTileM + r;
                if (slot_[sl] < n_slots) row_[sl] = A.m.slot_row ? (int64_t)A.m.slot_row[slot_[sl]] : slot_[sl];
                if (sl == 0) { mbar_wait(&f32_full[0], fph0); fph0 ^= 1; }
                else         { mbar_wait(&f32_full[1], fph1); fph1 ^= 1; }
            }
            for (int gi = 0; gi < n_gemm; ++gi) {
                const TcGemm& g = P.g[gi];
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl == 1 && !valid1) continue;
                    if (sl == 0) { mbar_wait(&acc_full[0], aph0); aph0 ^= 1; }
                    else         { mbar_wait(&acc_full[1], aph1); aph1 ^= 1; }
                    tc_fence_after();
                    const uint32_t t_acc = t_lane + (uint32_t)sl * 256u;
                    const float* Fb = F32 + (size_t)sl * (SL.f32_stride / 4);
                    const float* bias = Fb + g.bias_off;
                    const int64_t row = row_[sl], slot = slot_[sl];
                    if (g.epi == EPI_RGB) {
                        if (part == 0) {
                            uint32_t v[32];
                            tmem_ld32(t_acc, v);
                            tmem_ld_wait();
                            if (row >= 0) {
                                const NetDims& nd = A.m.nd;
                                const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                                const float w = A.m.slot_w ? A.m.slot_w[slot] : 1.0f;
                                const float sigma = sigma_[sl];
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
                        }
                    } else {
                        const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                        const bool publish = !(want_sigma && A.m.sigma_only);
                        const float* sw = Fb + P.sigma_w_off;
                        unsigned char* Hsl = Hs + (size_t)sl * h_bytes;
                        float sacc = 0.0f;
                        const int nslab = (g.n + 63) >> 6;
                        for (int j = 0; j < nslab; ++j) {
                            const int c0 = 64 * j + 16 * part;
                            if (c0 < g.n) {
                                unsigned char* dst = Hsl + (size_t)(c0 >> 3) * (kTileM * 16) + (size_t)r * 16;
                                if (g.epi == EPI_RELU)
                                    epi_piece16<false, true, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, 0, true);
                                else if (g.epi == EPI_LINEAR)
                                    epi_piece16<false, false, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, 0, true);
                                else
                                    sacc += epi_piece16<false, true, true>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, 0, publish);
                            }
                        }
                        if (publish) fence_proxy_async();
                        if (want_sigma) {
                            SIGP[part * kTileM + r] = sacc;
                            asm volatile("bar.sync 1, 512;" ::: "memory");
                            if (part == 0) {
                                float s = ((SIGP[r] + SIGP[kTileM + r]) + (SIGP[2 * kTileM + r] + SIGP[3 * kTileM + r])) + sw[L];
                                if (A.m.sigma_noise && row >= 0) s = s + A.m.sigma_noise[row];
                                const float sg = A.m.nd.softplus ? mn_softplus_shifted(s) : fmaxf(s, 0.0f);
                                sigma_[sl] = sg;
                                if (A.m.sigma_only && row >= 0) {
                                    const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                                    A.m.out[o] = A.m.slot_w ? sg * A.m.slot_w[slot] : sg;
                                }
                            }
                            asm volatile("bar.sync 1, 512;" ::: "memory");
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(epi_done_l + 8u * (uint32_t)sl);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&f32_empty[0]);
                if (valid1) mbar_arrive(&f32_empty[1]);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == kWarpProd) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

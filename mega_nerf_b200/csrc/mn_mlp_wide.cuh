// 512-wide sub-module MLP on tcgen05 (included inside mn_mlp_tc.cu's anonymous namespace).
//
// models/nerf.py:115-160 with layer_dim = 512 (the reference's 25 x 512 configurations): one 128-row tile per CTA
// iteration.  A 128 x 512 fp32 accumulator fills all 512 TMEM columns and the fp16 activations of one layer fill
// 128 KiB of shared memory, so neither can be double-buffered.  The kernel still keeps the tensor pipe busy by
// running every GEMM as two N = 256 halves, half 0 first:
//
//   MMA     : [H0 k<256][H0 k>=256][H1 k<256][H1 k>=256] | next layer [H0 k<256] ...
//   epilogue:                                  [epi H0   ][epi H1    ]
//
//  * epi H0 (drains TMEM columns 0..255, overwrites activation columns 0..255 IN PLACE) may start once H1 has
//    consumed activation columns 0..255, i.e. it runs under the second half of H1's MMAs;
//  * epi H1 runs under the next layer's [H0 k<256] MMAs, which only need activation columns 0..255 and TMEM
//    columns 0..255 - both released by epi H0;
//  * the next layer's [H0 k>=256] waits for epi H1.
//
// Weight images are half-major ([N half][K/8][256][8] fp16) so that a ring stage (32 K-columns x 256 outputs, 16 KiB)
// is one contiguous 1-D TMA copy.  Biases of the current GEMM and the sigma weights are staged by the epilogue warps
// themselves (plain loads from L2, double-buffered by GEMM parity) - the 24 KiB an all-GEMM bias block would take
// are needed for the weight ring.
//
// CS > 1: the kernel runs as thread-block clusters of CS CTAs whose tiles belong to the same sub-module (bucket
// alignment).  All CTAs of a cluster consume the identical weight stream, so each one fetches only 1/CS of every
// ring stage and MULTICASTS it into the shared memory of all CS CTAs (cp.async.bulk ... .multicast::cluster): the
// L2 -> SM weight traffic, which caps this kernel (~12.7 TB/s chip-wide at CS = 1), drops by CS.  A stage is
// refilled only after every CTA of the cluster has consumed it: tcgen05.commit multicasts the release to all CS
// `empty` barriers (count CS).
// Feature segments travel through the ring like in tc_mlp_pp_kernel: 16 K-columns of weights + the same 16 K-columns
// of the tile's feature image per stage (re-streamed for the second N half).
constexpr int kWMaxStages = 8;
constexpr int kWSlabCols = 32;
constexpr int kWHalf = 256;
constexpr int kWStageBytes = kWSlabCols * kWHalf * 2;

struct WLayout {
    int ring, h, bias, sw, sigp, bars, total, stages;
};

__host__ __device__ inline WLayout w_layout(const TcPlan& p) {
    WLayout s;
    const int fixed = p.L * kTileM * 2 + 2 * p.L * 4 + ((p.L + 4) * 4 + 15) / 16 * 16 + 2048 + 256;
    int st = (kSmemMax - fixed) / kWStageBytes;
    if (st > kWMaxStages) st = kWMaxStages;
    s.stages = st;
    s.ring = 0;
    s.h = st * kWStageBytes;
    s.bias = s.h + p.L * kTileM * 2;
    s.sw = s.bias + 2 * p.L * 4;
    s.sigp = s.sw + ((p.L + 4) * 4 + 15) / 16 * 16;
    s.bars = s.sigp + 2048;
    s.total = s.bars + 256;
    return s;
}

// mma_stage with the stage release multicast to the `empty` barrier of every CTA in the cluster
__device__ __forceinline__ void mma_stage_mc(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint64_t ad2, uint64_t bd2, uint32_t idesc,
                                             uint32_t accum, uint32_t two, uint32_t empty_bar_addr, uint32_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred e, p, q;\n\t.reg .b16 msk;\n\t"
        "cvt.u16.u32 msk, %9;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.and.b32 q, %7, 0, e;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], msk;\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(bd), "l"(ad2), "l"(bd2), "r"(idesc), "r"(accum), "r"(two), "r"(empty_bar_addr), "r"(cta_mask)
        : "memory");
}
// 1-D TMA copy global -> the same shared-memory offset in every CTA of `cta_mask`; each destination CTA's barrier
// (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint32_t cta_mask) {
    asm volatile(
        "{\n\t.reg .b16 msk;\n\tcvt.u16.u32 msk, %4;\n\t"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], msk;\n\t}"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "r"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_w() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS>
__global__ void __launch_bounds__(kThreads, 1) tc_mlp_wide_kernel(const TcArgs A) {   // CS > 1: launched with cluster dimension CS
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const WLayout SL = w_layout(P);
    unsigned char* ring = smem + SL.ring;
    unsigned char* Hs = smem + SL.h;
    const int kWStages = SL.stages;
    float* BIAS = reinterpret_cast<float*>(smem + SL.bias);   // [2][L]
    float* SW = reinterpret_cast<float*>(smem + SL.sw);       // sigma_w[L], sigma_b
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;             // [<=8]
    uint64_t* empty = bars + 8;        // [<=8]
    uint64_t* acc_ready = bars + 16;   // [2] per N half
    uint64_t* epi_done = bars + 18;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;
    const int L = P.L;
    // cluster c works on tile groups c, c + n_clusters, ...; CTA `rank` of the cluster takes tile group * CS + rank.
    // A CTA whose tile lies past the end still consumes the weight stream (its rows are all padding).
    uint32_t rank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t cta_mask = (1u << CS) - 1u;
    const int64_t n_groups = (n_tiles + CS - 1) / CS;
    const int64_t g_first = blockIdx.x / CS, g_stride = gridDim.x / CS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CS); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_ready[i], 1); mbar_init(&epi_done[i], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpProd) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    if (CS > 1) cluster_sync_w(); else __syncthreads();
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

    if (warp == kWarpProd) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, p_ahead = 0;
            const uint32_t empty_pa = smem_u32(empty);
            for (int64_t grp = g_first; grp < n_groups; grp += g_stride) {
                const int64_t tile_raw = grp * CS + rank;
                const int64_t tile = tile_raw < n_tiles ? tile_raw : grp * CS;   // padding CTA: any valid feature tile will do
                const unsigned char* wsub = A.wpack + (size_t)sub_of(grp * CS) * P.sub_bytes;
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int nw = g.n < kWHalf ? g.n : kWHalf;
                    const int nh = g.n / nw;
                    const int ktot = g.k[0] + (g.nseg > 1 ? g.k[1] : 0);
                    for (int half = 0; half < nh; ++half) {
                        const unsigned char* wimg = wsub + g.w_off + (size_t)half * ktot * nw * 2;
                        int kbase = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            if (g.src[sgi] != SRC_H) {
                                // feature segment: per stage 16 K-columns of weights (multicast slices when CS > 1) + this CTA's
                                // own 16 feature columns (4 KiB at +8 KiB)
                                const __half* xt = A.ximg + tile * (int64_t)(P.kpe + P.kaux) * kTileM +
                                                   (g.src[sgi] == SRC_XAUX ? (int64_t)P.kpe * kTileM : 0);
                                const uint32_t wbytes = (uint32_t)(kPPXCols * nw * 2), xbytes = (uint32_t)(kPPXCols * kTileM * 2);
                                for (int k0 = 0; k0 < kseg; k0 += kPPXCols) {
                                    const int cur = stage;
                                    unsigned char* st_base = ring + (size_t)cur * kWStageBytes;
                                    const unsigned char* wsrc = wimg + (size_t)(kbase + k0) * nw * 2;
                                    if (!p_ahead) mbar_wait(&empty[cur], phase ^ 1);
                                    if (++stage == kWStages) { stage = 0; phase ^= 1; }
                                    p_ahead = mbar_test_a(empty_pa + 8u * (uint32_t)stage, phase ^ 1);   // look-ahead probe (see mn_mlp_tc.cu)
                                    mbar_expect_tx(&full[cur], wbytes + xbytes);
                                    if (CS > 1) {
                                        const uint32_t slice = wbytes / CS;
                                        bulk_g2s_mc(st_base + (size_t)rank * slice, wsrc + (size_t)rank * slice, slice, &full[cur], cta_mask);
                                    } else {
                                        bulk_g2s(st_base, wsrc, wbytes, &full[cur]);
                                    }
                                    bulk_g2s(st_base + kPPXOff, xt + (size_t)k0 * kTileM, xbytes, &full[cur]);
                                }
                                kbase += kseg;
                                continue;
                            }
                            for (int k0 = 0; k0 < kseg; k0 += kWSlabCols) {
                                const int kc = min(kWSlabCols, kseg - k0);
                                const uint32_t bytes = (uint32_t)(kc * nw * 2);
                                const int cur = stage;
                                if (!p_ahead) mbar_wait(&empty[cur], phase ^ 1);      // CS > 1: released by every CTA of the cluster
                                if (++stage == kWStages) { stage = 0; phase ^= 1; }
                                p_ahead = mbar_test_a(empty_pa + 8u * (uint32_t)stage, phase ^ 1);
                                mbar_expect_tx(&full[cur], bytes);
                                if (CS > 1) {
                                    const uint32_t slice = bytes / CS;
                                    bulk_g2s_mc(ring + (size_t)cur * kWStageBytes + (size_t)rank * slice,
                                                wimg + (size_t)(kbase + k0) * nw * 2 + (size_t)rank * slice, slice, &full[cur], cta_mask);
                                } else {
                                    bulk_g2s(ring + (size_t)cur * kWStageBytes, wimg + (size_t)(kbase + k0) * nw * 2, bytes, &full[cur]);
                                }
                            }
                            kbase += kseg;
                        }
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer (whole warp, one elected lane issues) ===========================
        int stage = 0;
        uint32_t phase = 0, eph0 = 0, eph1 = 0, ahead = 0;
        bool pend0 = false, pend1 = false;   // an epilogue of that N half is outstanding (TMEM half + activation columns busy)
        const uint32_t h_base = smem_u32(Hs), ring_base = smem_u32(ring);
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const uint64_t xd0 = make_desc(ring_base + (uint32_t)kPPXOff, kTileM * 16, 128);
        const uint32_t acc_ready_a = smem_u32(acc_ready), epi_done_a = smem_u32(epi_done);
        const uint64_t a_step = (uint64_t)((2 * kTileM * 16) >> 4);
        const uint64_t st_step = (uint64_t)(kWStageBytes >> 4);
        for (int64_t grp = g_first; grp < n_groups; grp += g_stride) {
            for (int gi = 0; gi < n_gemm; ++gi) {
                const TcGemm& g = P.g[gi];
                const int nw = g.n < kWHalf ? g.n : kWHalf;
                const int nh = g.n / nw;
                const uint32_t idesc = make_idesc(nw);
                const uint64_t b_step = (uint64_t)((2 * nw * 16) >> 4);
                const uint64_t bd0 = make_desc(ring_base, (uint32_t)nw * 16, 128);
                const bool has_h = g.src[0] == SRC_H || (g.nseg > 1 && g.src[1] == SRC_H);
                for (int half = 0; half < nh; ++half) {
                    // TMEM columns of this half drained; for half 0 also: activation columns 0..255 of the previous GEMM written
                    if (half == 0) { if (pend0) { mbar_wait_a(epi_done_a, eph0); eph0 ^= 1; pend0 = false; } }
                    else           { if (pend1) { mbar_wait_a(epi_done_a + 8, eph1); eph1 ^= 1; pend1 = false; } }
                    tc_fence_after();
                    if (lane == 0) trace_ev(A.desc_swap, 0, 1, half, gi);
                    const uint32_t d_tmem = tmem_base + (uint32_t)half * (uint32_t)kWHalf;
                    uint32_t accum = 0;
                    for (int sgi = 0; sgi < g.nseg; ++sgi) {
                        const bool from_x = g.src[sgi] != SRC_H;
                        if (from_x) {
                            for (int rem = g.k[sgi]; rem > 0; rem -= kPPXCols) {
                                const uint64_t so = (uint64_t)stage * st_step;
                                const uint32_t cur = (uint32_t)stage;
                                if (!ahead) mbar_wait_a(full_a + 8u * cur, phase);
                                tc_fence_after();
                                if (++stage == kWStages) { stage = 0; phase ^= 1; }
                                ahead = mbar_test_a(full_a + 8u * (uint32_t)stage, phase);      // look-ahead probe (see mn_mlp_tc.cu)
                                if (CS > 1)
                                    mma_stage_mc(d_tmem, xd0 + so, bd0 + so, 0, 0, idesc, accum, 0u, empty_a + 8u * cur, cta_mask);
                                else
                                    mma_stage(d_tmem, xd0 + so, bd0 + so, 0, 0, idesc, accum, 0u, empty_a + 8u * cur);
                                accum = 1;
                            }
                            continue;
                        }
                        uint64_t ad = make_desc(h_base, kTileM * 16, 128);
                        int kdone = 0;
                        for (int rem = g.k[sgi]; rem > 0; rem -= kWSlabCols) {
                            if (!from_x && kdone == kWHalf && pend1) {
                                // activation columns >= 256 come from the previous GEMM's half-1 epilogue
                                mbar_wait_a(epi_done_a + 8, eph1);
                                eph1 ^= 1;
                                pend1 = false;
                                tc_fence_after();
                            }
                            const uint32_t two = rem >= 32 ? 1u : 0u;
                            const uint64_t bd = bd0 + (uint64_t)stage * st_step;
                            const uint32_t cur = (uint32_t)stage;
                            if (!ahead) mbar_wait_a(full_a + 8u * cur, phase);
                            tc_fence_after();
                            if (++stage == kWStages) { stage = 0; phase ^= 1; }
                            ahead = mbar_test_a(full_a + 8u * (uint32_t)stage, phase);
                            if (CS > 1)
                                mma_stage_mc(d_tmem, ad, bd, ad + a_step, bd + b_step, idesc, accum, two, empty_a + 8u * cur, cta_mask);
                            else
                                mma_stage(d_tmem, ad, bd, ad + a_step, bd + b_step, idesc, accum, two, empty_a + 8u * cur);
                            accum = 1;
                            ad += two ? 2 * a_step : a_step;
                            kdone += kWSlabCols;
                            // half 0's accumulator is complete and nothing issued later reads activation columns 0..255:
                            // its epilogue may overwrite them while the rest of half 1 is still running
                            if (nh == 2 && half == 1 && !from_x && (kdone == kWHalf || (rem <= kWSlabCols && kdone < kWHalf)))
                                commit_elect(acc_ready_a);
                        }
                    }
                    if (nh == 2 && half == 0 && !has_h) commit_elect(acc_ready_a);        // feature-only GEMM (layer 0)
                    if (half == nh - 1) commit_elect(acc_ready_a + 8u * (uint32_t)half);  // last half: whole GEMM issued
                    if (lane == 0) trace_ev(A.desc_swap, 0, 2, half, gi);
                }
                // a pending half-1 epilogue that this GEMM never had to wait for is still outstanding; it is
                // collected before TMEM half 1 / activation columns >= 256 are touched again
                pend0 = true;
                if (nh == 2) pend1 = true;
            }
        }
    } else {
        // =========================== epilogue (16 warps) ===========================
        const int q = warp & 3;
        const int part = warp >> 2;
        const int r = q * 32 + lane;
        const int etid = threadIdx.x;   // 0..511
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t aph0 = 0, aph1 = 0, gidx = 0;
        for (int64_t grp = g_first; grp < n_groups; grp += g_stride) {
            const int64_t tile = grp * CS + rank;
            const int64_t slot = tile * kTileM + r;
            int64_t row = -1;
            if (slot < n_slots) row = A.m.slot_row ? (int64_t)A.m.slot_row[slot] : slot;
            const int sub = sub_of(grp * CS);
            const float* fb = reinterpret_cast<const float*>(A.wpack + (size_t)sub * P.sub_bytes + (size_t)P.plane_bytes * 2);
            // sigma weights of this tile's sub-module (read only by the last trunk layer's epilogue; the bar.sync
            // of the first GEMM below orders these stores before any read)
            for (int i = etid; i < L + 4; i += kEpiWarps * 32) SW[i] = __ldg(fb + P.sigma_w_off + i);
            float sigma = 0.0f;
            for (int gi = 0; gi < n_gemm; ++gi, ++gidx) {
                const TcGemm& g = P.g[gi];
                const int nw = g.n < kWHalf ? g.n : kWHalf;
                const int nh = g.n / nw;
                float* bias = BIAS + (size_t)(gidx & 1u) * L;   // consecutive GEMMs (also across tiles) alternate buffers
                for (int i = etid; i < g.n; i += kEpiWarps * 32) bias[i] = __ldg(fb + g.bias_off + i);
                asm volatile("bar.sync 1, 512;" ::: "memory");
                float sacc = 0.0f;
                for (int half = 0; half < nh; ++half) {
                    if (half == 0) { mbar_wait(&acc_ready[0], aph0); aph0 ^= 1; }
                    else           { mbar_wait(&acc_ready[1], aph1); aph1 ^= 1; }
                    tc_fence_after();
                    if (warp == 0 && lane == 0) trace_ev(A.desc_swap, 1, 3, half, gi);
                    const uint32_t t_acc = t_lane + (uint32_t)half * (uint32_t)kWHalf;
                    if (g.epi == EPI_RGB) {
                        if (part == 0) {
                            uint32_t v[32];
                            tmem_ld32(t_acc, v);
                            tmem_ld_wait();
                            if (row >= 0) tc_emit_rgb(A.m, sub, row, slot, v, bias, sigma);
                        }
                    } else {
                        const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                        const bool publish = !(want_sigma && A.m.sigma_only);
                        for (int j = 0; j < 4; ++j) {
                            const int c0 = 64 * j + 16 * part;
                            if (c0 < nw) {
                                const int col = half * kWHalf + c0;
                                unsigned char* dst = Hs + (size_t)(col >> 3) * (kTileM * 16) + (size_t)r * 16;
                                if (g.epi == EPI_RELU)
                                    epi_piece16<false, true, false>(t_acc + (uint32_t)c0, bias + col, SW + col, dst, 0, true);
                                else if (g.epi == EPI_LINEAR)
                                    epi_piece16<false, false, false>(t_acc + (uint32_t)c0, bias + col, SW + col, dst, 0, true);
                                else
                                    sacc += epi_piece16<false, true, true>(t_acc + (uint32_t)c0, bias + col, SW + col, dst, 0, publish);
                            }
                        }
                        if (publish) fence_proxy_async();
                        if (want_sigma && half == nh - 1) {
                            SIGP[part * kTileM + r] = sacc;
                            asm volatile("bar.sync 1, 512;" ::: "memory");
                            if (part == 0) {
                                float s = ((SIGP[r] + SIGP[kTileM + r]) + (SIGP[2 * kTileM + r] + SIGP[3 * kTileM + r])) + SW[L];
                                if (A.m.sigma_noise && row >= 0) s = s + A.m.sigma_noise[row];
                                sigma = A.m.nd.softplus ? mn_softplus_shifted(s) : fmaxf(s, 0.0f);
                                if (A.m.sigma_only && row >= 0) {
                                    const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                                    A.m.out[o] = A.m.slot_w ? sigma * A.m.slot_w[slot] : sigma;
                                }
                            }
                            asm volatile("bar.sync 1, 512;" ::: "memory");   // SIGP / SW are rewritten by the next tile
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (warp == 0 && lane == 0) trace_ev(A.desc_swap, 1, 4, half, gi);
                    if (lane == 0) mbar_arrive(&epi_done[half]);
                }
            }
        }
    }
    tc_fence_before();
    if (CS > 1) cluster_sync_w(); else __syncthreads();   // peers may still multicast into / arrive on this CTA's shared memory
    if (warp == kWarpProd) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// CTA-pair variant of the ping-pong MLP kernel (included inside mn_mlp_tc.cu's anonymous namespace).
//
// A cluster of two CTAs (one SM pair) runs tcgen05.mma.cta_group::2: one instruction issued by the leader CTA drives
// both SMs' tensor cores on a 256-row tile (128 rows per CTA).  Each CTA stages only HALF of every weight slab
// (N/2 rows of B): measured on B200, the single-CTA kernel sits on the L2 -> SM delivery ceiling (~32 B/clk/SM: each
// 128-row tile re-reads 1.2 MB of weights) and its tensor core fetches 12 KiB of operands per 128x256x16
// MMA (A 4 KiB + B 8 KiB, ~182 cycles instead of 128); the pair halves the weight bytes delivered to each SM and the B
// operand bytes read per SM, and a ring stage holds 64 K-columns (4 MMAs).
// Everything else follows tc_mlp_pp_kernel: two 256-row tiles (X, Y) per cluster - four consecutive 128-row tiles,
// one sub-module thanks to the 512-row bucket alignment - GEMMs issued X_l, Y_l, X_l+1, ..., each CTA's 16 epilogue
// warps drain its own 128 TMEM lanes under the other tile's MMAs; feature columns ride in the weight ring.
//
// Cross-CTA signalling (all barriers live at identical offsets in both CTAs):
//   full[s]            the LEADER's barrier: its producer arrives with expect_tx for the bytes of both CTAs, and both
//                      CTAs' tensor-map TMA copies (.cta_group::2) complete_tx on it (shared::cluster address from mapa);
//   empty[s], acc_full[slot]   tcgen05.commit ... multicast::cluster to both CTAs;
//   epi_done[slot]     leader's barrier, 32 arrivals (16 epilogue warps per CTA, the peer's arrive remotely).
//                      With A.c2_relay (MN_TC_C2=2) the peer's warps arrive on its own epi_local[slot] instead and the
//                      peer's otherwise idle MMA warp forwards one remote arrive: 16 local + 1 remote arrivals.
#pragma once


constexpr int kC2MaxStages = 16;
// Ring geometry.  Streaming mode: a stage = 64 K-columns x (N/2 <= 128) weight rows (<= 16 KiB), or, for a feature
// segment, 32 K-columns of weights (<= 8 KiB) + the same 32 feature columns of the CTA's tile (8 KiB at +8 KiB).
// Shared mode (TcArgs::c2_share): half-size stages - 32 K-columns (8 KiB) / 16 + 16 (4 KiB + 4 KiB) - so that the 8 stages
// a 256-deep layer pins while both tile slots consume them leave room for the next layer's first stages.
struct C2Layout {
    int ring, h, f32, f32_stride, sigp, bars, total, stages;
    int stage_cols, stage_bytes, xcols, xoff;
};

__host__ __device__ inline C2Layout c2_layout(const TcPlan& p, bool share) {
    C2Layout s;
    s.stage_cols = share ? 32 : 64;
    s.stage_bytes = s.stage_cols * 128 * 2;
    s.xcols = s.stage_cols / 2;
    s.xoff = s.stage_bytes / 2;
    s.f32_stride = ((p.f32_floats * 4 + 15) / 16) * 16;
    const int fixed = 2 * p.L * kTileM * 2 + s.f32_stride + 2048 + 512;
    int st = (kSmemMax - fixed) / s.stage_bytes;
    if (st > kC2MaxStages) st = kC2MaxStages;
    s.stages = st;
    s.ring = 0;
    s.h = st * s.stage_bytes;
    s.f32 = s.h + 2 * p.L * kTileM * 2;
    s.sigp = s.f32 + s.f32_stride;
    s.bars = s.sigp + 2048;
    s.total = s.bars + 512;
    return s;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
// Remote arrive with the DEFAULT semantics (release at CTA scope), the form CUTLASS's ClusterBarrier::arrive(cta_id) uses.
// `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive - a GPU-scope barrier on the critical
// path of every epilogue (~1 us under load) - and `.acquire.cluster` waits to a CCTL.IVALL (L1 invalidate) after every
// successful try_wait of the MMA warp.  Neither is needed here: what crosses the CTA boundary is shared memory written
// through the generic proxy and already fenced to the async proxy by its writers (fence.proxy.async), read by
// tcgen05.mma; nothing in global memory is published through these barriers.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
// (release == 0: the stage stays resident - the other tile slot consumes it next, shared mode)
__device__ __forceinline__ void c2_stage(uint32_t d_tmem, uint64_t ad, uint64_t a_step, uint64_t bd, uint64_t b_step, uint32_t idesc,
                                         uint32_t accum, int nk, uint32_t empty_bar, uint32_t release = 1u) {
    asm volatile(
        "{\n\t.reg .pred e, p, q1, q2, q3, r;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t.reg .b16 msk;\n\t"
        "mov.b16 msk, 3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.gt.and.s32 q1, %7, 1, e;\n\t"
        "setp.gt.and.s32 q2, %7, 2, e;\n\t"
        "setp.gt.and.s32 q3, %7, 3, e;\n\t"
        "setp.ne.and.b32 r, %9, 0, e;\n\t"
        "add.u64 a1, %1, %2;\n\tadd.u64 a2, a1, %2;\n\tadd.u64 a3, a2, %2;\n\t"
        "add.u64 b1, %3, %4;\n\tadd.u64 b2, b1, %4;\n\tadd.u64 b3, b2, %4;\n\t"
        "@e  tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %3, %5, p;\n\t"
        "@q1 tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %5, 1;\n\t"
        "@q2 tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %5, 1;\n\t"
        "@q3 tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %5, 1;\n\t"
        "@r  tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], msk;\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(a_step), "l"(bd), "l"(b_step), "r"(idesc), "r"(accum), "r"(nk), "r"(empty_bar), "r"(release)
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

// First half of epi_piece16 for the trailing variant: TMEM -> +bias -> (ReLU) -> fp16 pairs kept in 8 registers (the
// stores happen later, slab by slab); returns the partial sigma dot product if kSigma.  Same arithmetic as epi_piece16.
template <bool kRelu, bool kSigma>
__device__ __forceinline__ float c2_convert16(uint32_t taddr, const float* __restrict__ bias16, const float* __restrict__ sw16,
                                              uint32_t* hpk) {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    const float4* b4 = reinterpret_cast<const float4*>(bias16);
    const float4 b0 = b4[0], b1 = b4[1], b2 = b4[2], b3 = b4[3];
    const float b[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
    tmem_ld_wait();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + b[i];
    if (kRelu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
    }
    float sacc = 0.0f;
    if (kSigma) {
        const float4* s4 = reinterpret_cast<const float4*>(sw16);
        const float4 s0 = s4[0], s1 = s4[1], s2 = s4[2], s3 = s4[3];
        const float sw[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc = fmaf(f[i], sw[i], sacc);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) hpk[e] = pack_h2(f[2 * e], f[2 * e + 1]);
    return sacc;
}

struct C2Maps {
    CUtensorMap w64, w32, w8, w4;   // packed weights viewed as rows of 256 B: boxes of 64 / 32 / 8 / 4 rows
    CUtensorMap x32, x16;       // feature tile images, same view: boxes of 32 / 16 rows (= K-columns)
};

// copies `rows` x 256 B starting at row `row0` of a rows-of-256-B tensor map into this CTA's shared memory, using the
// largest boxes first (box heights ra > rb > rc; a map pointer may be null)
__device__ __forceinline__ void c2_copy_rows(uint32_t dst, int row0, int rows, uint32_t bar, const CUtensorMap* ma, int ra,
                                             const CUtensorMap* mb, int rb, const CUtensorMap* mc, int rc,
                                             const CUtensorMap* md = nullptr, int rd = 0) {
    while (ma && rows >= ra) { tma2d_c2(dst, ma, row0, bar); dst += (uint32_t)ra * 256u; row0 += ra; rows -= ra; }
    while (mb && rows >= rb) { tma2d_c2(dst, mb, row0, bar); dst += (uint32_t)rb * 256u; row0 += rb; rows -= rb; }
    while (mc && rows >= rc) { tma2d_c2(dst, mc, row0, bar); dst += (uint32_t)rc * 256u; row0 += rc; rows -= rc; }
    while (md && rows >= rd) { tma2d_c2(dst, md, row0, bar); dst += (uint32_t)rd * 256u; row0 += rd; rows -= rd; }
}

// kTrail (MN_TC_C2=3): every epilogue warp first loads ALL its accumulator columns into registers, then converts and
// stores them slab by slab (64 columns), announcing each slab on h_ready[slot][slab]; the leader starts the slot's next
// GEMM as soon as slab 0 is announced by everybody (=> the whole accumulator has been read and may be overwritten) and
// issues the K-steps of slab j right after slab j - the chain MMA -> epilogue -> MMA of a slot shrinks from the full
// drain to one TMEM load + one slab.  Always uses the relay (one remote arrive per slab from the peer CTA).
template <bool kTrail>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    tc_mlp_c2_kernel(const TcArgs A, const __grid_constant__ C2Maps TM) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const bool share = A.c2_share != 0;
    const C2Layout SL = c2_layout(P, share);
    const int kStages = SL.stages;
    const int kC2StageCols = SL.stage_cols, kC2StageBytes = SL.stage_bytes, kC2XCols = SL.xcols, kC2XOff = SL.xoff;
    unsigned char* ring = smem + SL.ring;
    unsigned char* Hs = smem + SL.h;
    float* F32 = reinterpret_cast<float*>(smem + SL.f32);
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;            // [<=16] LEADER's: both CTAs' halves of the stage have landed
    uint64_t* empty = bars + 16;      // [<=16] per CTA (released by the leader's multicast commit)
    uint64_t* acc_full = bars + 32;   // [2]   per CTA
    uint64_t* epi_done = bars + 34;   // [2]   leader's
    uint64_t* f32_full = bars + 36;   //       per CTA
    uint64_t* f32_empty = bars + 37;  //       per CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);
    uint64_t* epi_local = bars + 40;  // [2]   per CTA, relay mode only
    uint64_t* h_ready = bars + 44;    // [2][4] leader's (kTrail): 16 local arrivals + 1 relayed from the peer
    uint64_t* h_local = bars + 52;    // [2][4] per CTA  (kTrail): the peer's 16 epilogue warps
    const bool relay = kTrail || A.c2_relay != 0;

    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;        // 128-row tiles
    const int64_t n_quads = (n_tiles + 3) / 4;                       // a cluster iteration = 4 consecutive tiles
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;
    const int h_bytes = P.L * kTileM * 2;
    const int64_t cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;        // cluster index / count
    const int xrows_tile = P.kpe + P.kaux;                           // 256-byte rows per feature tile image (one per K-column)

    if (threadIdx.x == 0) {
        for (int i = 0; i < kC2MaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&epi_done[i], relay ? kEpiWarps + 1 : 2 * kEpiWarps);
            mbar_init(&epi_local[i], kEpiWarps);
        }
        for (int i = 0; i < 8; ++i) { mbar_init(&h_ready[i], kEpiWarps + 1); mbar_init(&h_local[i], kEpiWarps); }
        mbar_init(f32_full, 1);
        mbar_init(f32_empty, kEpiWarps);
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
    clk_stamp(A.desc_swap, 0);

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
            uint32_t phase = 0, fph_e = 0, p_ahead = 0;
            int last_sub = -1;
            const uint32_t f32_bytes = (uint32_t)SL.f32_stride;
            const uint32_t empty_pa = smem_u32(empty);
            const uint32_t full_l = mapa_u32(smem_u32(full), 0);          // the LEADER's full[] barriers count both halves
            const uint32_t ring_a = smem_u32(ring);
            for (int64_t q = cl; q < n_quads; q += ncl) {
                const int sub0 = sub_of(4 * q);
                const unsigned char* wsub = A.wpack + (size_t)sub0 * P.sub_bytes;
                if (sub0 != last_sub) {           // bias / sigma block: re-staged only when the sub-module changes
                    if (last_sub >= 0) { mbar_wait(f32_empty, fph_e); fph_e ^= 1; }
                    mbar_expect_tx(f32_full, f32_bytes);
                    bulk_g2s(reinterpret_cast<unsigned char*>(F32), wsub + (size_t)P.plane_bytes * 2, f32_bytes, f32_full);
                    last_sub = sub0;
                }
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int nhalf = g.n >> 1;                                   // weight rows held by this CTA
                    const int K = g.k[0] + (g.nseg > 1 ? g.k[1] : 0);
                    const unsigned char* wimg = wsub + c2_off + g.w_off + (size_t)rank * K * nhalf * 2;
                    // shared mode: a GEMM whose only operand is the activation buffer (no per-tile feature columns in its
                    // stream) is loaded ONCE per cluster iteration; both tile slots consume the same stages
                    const bool shared_g = share && g.nseg == 1 && g.src[0] == SRC_H;
                    for (int sl = 0; sl < 2; ++sl) {
                        if (shared_g && sl == 1) break;
                        const int64_t my_tile = 4 * q + 2 * sl + rank;           // this CTA's 128-row tile of the slot
                        int kbase = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            const bool from_x = g.src[sgi] != SRC_H;
                            const int step = from_x ? kC2XCols : kC2StageCols;
                            const int xrow0 = (int)(my_tile * xrows_tile) + (g.src[sgi] == SRC_XAUX ? P.kpe : 0);
                            for (int k0 = 0; k0 < kseg; k0 += step) {
                                const int kc = min(step, kseg - k0);
                                const uint32_t wbytes = (uint32_t)(kc * nhalf * 2);
                                const uint32_t xbytes = from_x ? (uint32_t)(kc * kTileM * 2) : 0u;
                                if (!p_ahead) mbar_wait(&empty[stage], phase ^ 1);
                                const int cur = stage;
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                                p_ahead = mbar_test_a(empty_pa + 8u * (uint32_t)stage, phase ^ 1);    // look-ahead, overlaps the copies below
                                const uint32_t bar = full_l + 8u * (uint32_t)cur;
                                const uint32_t dst = ring_a + (uint32_t)cur * kC2StageBytes;
                                // the leader announces the bytes of BOTH CTAs (the streams are symmetric); the peer's copies
                                // only complete_tx on the leader's barrier - no remote arrive on the critical path
                                if (leader) mbar_expect_tx(&full[cur], 2u * (wbytes + xbytes));
                                const int wrow0 = (int)((size_t)((wimg + (size_t)(kbase + k0) * nhalf * 2) - A.wpack) >> 8);
                                c2_copy_rows(dst, wrow0, (int)(wbytes >> 8), bar, &TM.w64, 64, &TM.w32, 32, &TM.w8, 8, &TM.w4, 4);
                                if (from_x) c2_copy_rows(dst + kC2XOff, xrow0 + k0, kc, bar, &TM.x32, 32, &TM.x16, 16, nullptr, 0);
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
            uint32_t phase = 0, eph0 = 0, eph1 = 0, ahead = 0;
            bool started0 = false, started1 = false;
            const uint32_t h_base = smem_u32(Hs), ring_base = smem_u32(ring);
            const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
            const uint32_t acc_full_a = smem_u32(acc_full), epi_done_a = smem_u32(epi_done);
            const uint64_t a_step = (uint64_t)((2 * kTileM * 16) >> 4);
            const uint64_t st_step = (uint64_t)(kC2StageBytes >> 4);
            const uint64_t xd0 = make_desc(ring_base + (uint32_t)kC2XOff, kTileM * 16, 128);
            for (int64_t q = cl; q < n_quads; q += ncl) {
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int nhalf = g.n >> 1;
                    const uint32_t idesc = make_idesc_m(256, g.n);
                    const uint64_t b_step = (uint64_t)((2 * nhalf * 16) >> 4);
                    const uint64_t bd0 = make_desc(ring_base, (uint32_t)nhalf * 16, 128);
                    // shared mode: slot 0 walks the GEMM's stages WITHOUT releasing them, slot 1 walks the same stages again
                    // (their full barriers have completed: the parity waits pass immediately) and releases each one
                    const bool shared_g = share && g.nseg == 1 && g.src[0] == SRC_H;
                    const int stage_g0 = stage;
                    const uint32_t phase_g0 = phase;
                    for (int sl = 0; sl < 2; ++sl) {
                        if (shared_g && sl == 1) { stage = stage_g0; phase = phase_g0; ahead = 0; }
                        const uint32_t release = (shared_g && sl == 0) ? 0u : 1u;
                        const bool prev = sl == 0 ? started0 : started1;       // this slot has an epilogue in flight
                        const uint32_t hpar = sl == 0 ? eph0 : eph1;
                        const uint32_t hbar = smem_u32(h_ready) + 32u * (uint32_t)sl;
                        int waited = 0;                                         // kTrail: slab barriers of that epilogue already seen
                        if (kTrail) {
                            // slab 0 announced by all 32 warps => every accumulator column has been loaded: it may be overwritten
                            if (prev) { mbar_wait_cluster(hbar, hpar); waited = 1; }
                        } else {
                            // both CTAs have drained this slot's accumulator and written its activations
                            if (prev) mbar_wait_cluster(epi_done_a + 8u * (uint32_t)sl, hpar);
                        }
                        if (prev && !kTrail) { if (sl == 0) eph0 ^= 1; else eph1 ^= 1; }
                        if (sl == 0) started0 = true; else started1 = true;
                        tc_fence_after();
                        if (lane == 0) trace_ev(A.desc_swap, 0, 1, sl, gi);
                        const uint32_t d_tmem = tmem_base + (uint32_t)sl * 256u;
                        uint32_t accum = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            const bool from_x = g.src[sgi] != SRC_H;
                            const int step = from_x ? kC2XCols : kC2StageCols;
                            uint64_t ad = make_desc(h_base + (uint32_t)(sl * h_bytes), kTileM * 16, 128);
                            for (int k0 = 0; k0 < kseg; k0 += step) {
                                const int kc = min(step, kseg - k0);
                                const uint64_t so = (uint64_t)stage * st_step;
                                if (kTrail && prev && !from_x) {
                                    // this stage reads activation columns [k0, k0 + 64) = slab k0 / 64 of the previous epilogue
                                    const int hs = k0 >> 6;
                                    while (waited <= hs) { mbar_wait_cluster(hbar + 8u * (uint32_t)waited, hpar); ++waited; }
                                }
                                const uint32_t cur = (uint32_t)stage;
                                if (!ahead) mbar_wait_cluster(full_a + 8u * cur, phase);
                                tc_fence_after();
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                                // look-ahead probe of the next stage (streaming mode only: in shared mode slot 1 rewinds), issued
                                // before this stage's MMAs so that its latency overlaps them
                                ahead = share ? 0u : mbar_test_a(full_a + 8u * (uint32_t)stage, phase);
                                c2_stage(d_tmem, from_x ? xd0 + so : ad, a_step, bd0 + so, b_step, idesc, accum, kc >> 4,
                                         empty_a + 8u * cur, release);
                                accum = 1;
                                ad += (uint64_t)(kc >> 4) * a_step;
                            }
                        }
                        if (kTrail && prev) {
                            // every slab barrier is consumed exactly once per GEMM (parity waits must not skip a phase)
                            while (waited < 4) { mbar_wait_cluster(hbar + 8u * (uint32_t)waited, hpar); ++waited; }
                            if (sl == 0) eph0 ^= 1; else eph1 ^= 1;
                        }
                        c2_commit_both(acc_full_a + 8u * (uint32_t)sl);
                        if (lane == 0) trace_ev(A.desc_swap, 0, 2, sl, gi);
                    }
                }
            }
        } else if (relay && lane == 0) {
            // peer CTA, relay mode: one remote arrive per (GEMM, slot) once all 16 local epilogue warps are done.
            // (their release.cta arrives -> this acquire -> the release.cluster arrive below: the activation writes are
            // ordered before the leader's acquire.cluster wait; they were already fenced to the async proxy by their writers)
            uint32_t lph0 = 0, lph1 = 0;
            const uint32_t epi_done_l = mapa_u32(smem_u32(epi_done), 0);
            const uint32_t h_ready_l = mapa_u32(smem_u32(h_ready), 0);
            for (int64_t q = cl; q < n_quads; q += ncl)
                for (int gi = 0; gi < n_gemm; ++gi)
                    for (int sl = 0; sl < 2; ++sl) {
                        const uint32_t lp = sl == 0 ? lph0 : lph1;
                        if (kTrail) {
                            for (int j = 0; j < 4; ++j) {
                                mbar_wait(&h_local[sl * 4 + j], lp);
                                mbar_arrive_cluster(h_ready_l + 8u * (uint32_t)(sl * 4 + j));
                            }
                        } else {
                            mbar_wait(&epi_local[sl], lp);
                            mbar_arrive_cluster(epi_done_l + 8u * (uint32_t)sl);
                        }
                        if (sl == 0) lph0 ^= 1; else lph1 ^= 1;
                    }
        }
    } else {
        // =========================== epilogue (16 warps per CTA, own 128 TMEM lanes) ===========================
        const int q4 = warp & 3;
        const int part = warp >> 2;
        const int r = q4 * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
        uint32_t aph0 = 0, aph1 = 0, fph = 0;
        int last_sub = -1;
        const int L = P.L;
        const uint32_t epi_done_l = mapa_u32(smem_u32(epi_done), 0);      // leader's barrier (local address if we are the leader)
        for (int64_t q = cl; q < n_quads; q += ncl) {
            int64_t slot_[2], row_[2] = {-1, -1};
            float sigma_[2] = {0.0f, 0.0f};
            for (int sl = 0; sl < 2; ++sl) {
                slot_[sl] = (4 * q + 2 * sl + rank) * kTileM + r;
                if (slot_[sl] < n_slots) row_[sl] = A.m.slot_row ? (int64_t)A.m.slot_row[slot_[sl]] : slot_[sl];
            }
            const int sub = sub_of(4 * q);
            if (sub != last_sub) {
                if (last_sub >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(f32_empty);
                }
                mbar_wait(f32_full, fph);
                fph ^= 1;
                last_sub = sub;
            }
            for (int gi = 0; gi < n_gemm; ++gi) {
                const TcGemm& g = P.g[gi];
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl == 0) { mbar_wait(&acc_full[0], aph0); aph0 ^= 1; }
                    else         { mbar_wait(&acc_full[1], aph1); aph1 ^= 1; }
                    tc_fence_after();
                    if (warp == 0 && lane == 0) trace_ev(A.desc_swap, 1, 3, sl, gi);
                    const uint32_t t_acc = t_lane + (uint32_t)sl * 256u;
                    const float* bias = F32 + g.bias_off;
                    const int64_t row = row_[sl], slot = slot_[sl];
                    // kTrail: announce slab j of this slot (leader: own barrier; peer: local barrier, forwarded by the relay)
                    auto announce = [&](int j) {
                        __syncwarp();
                        if (lane == 0) {
                            if (leader) mbar_arrive(&h_ready[sl * 4 + j]);
                            else mbar_arrive(&h_local[sl * 4 + j]);
                        }
                    };
                    if (g.epi == EPI_RGB) {
                        uint32_t v[32];
                        if (part == 0) {
                            tmem_ld32(t_acc, v);
                            tmem_ld_wait();
                        }
                        if (kTrail) {
                            tc_fence_before();
                            for (int j = 0; j < 4; ++j) announce(j);
                        }
                        if (part == 0 && row >= 0) tc_emit_rgb(A.m, sub, row, slot, v, bias, sigma_[sl]);
                    } else if (kTrail) {
                        const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                        const bool publish = !(want_sigma && A.m.sigma_only);
                        const float* sw = F32 + P.sigma_w_off;
                        unsigned char* Hsl = Hs + (size_t)sl * h_bytes;
                        float sacc = 0.0f;
                        const int nslab = (g.n + 63) >> 6;
                        // the warp's four 16-column pieces as fp16 pairs (named arrays: kept in registers)
                        uint32_t hp0[8], hp1[8], hp2[8], hp3[8];
#define MN_C2_CONVERT(J, HP)                                                                                              \
    {                                                                                                                    \
        const int c0 = 64 * (J) + 16 * part;                                                                             \
        if ((J) < nslab && c0 < g.n) {                                                                                   \
            if (g.epi == EPI_RELU) c2_convert16<true, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, HP);              \
            else if (g.epi == EPI_LINEAR) c2_convert16<false, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, HP);      \
            else sacc += c2_convert16<true, true>(t_acc + (uint32_t)c0, bias + c0, sw + c0, HP);                         \
        }                                                                                                                \
    }
                        MN_C2_CONVERT(0, hp0) MN_C2_CONVERT(1, hp1) MN_C2_CONVERT(2, hp2) MN_C2_CONVERT(3, hp3)
#undef MN_C2_CONVERT
                        tc_fence_before();      // all of this warp's TMEM reads are done: slab 0's announcement releases the accumulator
#define MN_C2_STORE(J, HP)                                                                                               \
    {                                                                                                                    \
        const int c0 = 64 * (J) + 16 * part;                                                                             \
        if ((J) < nslab && c0 < g.n && publish) {                                                                        \
            unsigned char* dst = Hsl + (size_t)(c0 >> 3) * (kTileM * 16) + (size_t)r * 16;                               \
            *reinterpret_cast<uint4*>(dst) = make_uint4(HP[0], HP[1], HP[2], HP[3]);                                     \
            *reinterpret_cast<uint4*>(dst + kTileM * 16) = make_uint4(HP[4], HP[5], HP[6], HP[7]);                       \
            fence_proxy_async();                                                                                         \
        }                                                                                                                \
        announce(J);                                                                                                     \
    }
                        MN_C2_STORE(0, hp0) MN_C2_STORE(1, hp1) MN_C2_STORE(2, hp2) MN_C2_STORE(3, hp3)
#undef MN_C2_STORE
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
                    } else {
                        const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                        const bool publish = !(want_sigma && A.m.sigma_only);
                        const float* sw = F32 + P.sigma_w_off;
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
                    if (warp == 0 && lane == 0) trace_ev(A.desc_swap, 1, 4, sl, gi);
                    if (lane == 0 && !kTrail) {
                        if (relay && !leader) mbar_arrive(&epi_local[sl]);
                        else mbar_arrive_cluster(epi_done_l + 8u * (uint32_t)sl);
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    clk_stamp(A.desc_swap, 1);
    if (warp == kWarpProd) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// TMEM ping-pong kernel: the inference MLP for layer_dim <= 256 (included inside mn_mlp_tc.cu's anonymous namespace).
//
// Shared-memory bandwidth bounds tc_mlp_pp_kernel: per 128 x 256 x 256 layer of ONE tile it moves 64 KiB (A read) + 128 KiB
// (B read) + 128 KiB (TMA weight fill) + 64 KiB (epilogue stores) through a 128 B/clk port, i.e. 3072 clk for 2048 clk of
// MMA (ncu: shared pipe 92 % busy at 0.6 of the tensor peak).  This kernel removes the two activation terms and halves the fill:
//   * the A operand of every hidden layer lives in TENSOR MEMORY (tcgen05.mma with A from TMEM); the epilogue writes it
//     there with tcgen05.st (fp16 pairs, one TMEM lane per row) - activations never touch shared memory;
//   * a CTA works on a PAIR of tiles X, Y of one sub-module and every ring stage (16 KiB: 64 K-columns x one 128-wide
//     N-half of the weights) is fetched ONCE and multiplied into both tiles;
//   * every layer is issued as two N = 128 halves in the order X.h0, Y.h0, X.h1, Y.h1.  Each block is 1024 clk of MMA
//     at peak; the epilogue of a block runs under the next block (the other tile), the one write that sits on the
//     dependent chain (h1 -> next layer's A) has the other tile's h1 block to hide under.  The h0 half of a layer's
//     output waits in REGISTERS until the layer's last MMA has read the old A, then both halves are stored.
//   TMEM columns per tile slot s: [256 s, +128) accumulator (one N-half), [256 s + 128, +128) A operand (K <= 256 fp16).
//   B read 128 KiB + fill 64 KiB per tile-layer = 1536 clk of the shared pipe for 2048 clk of MMA (ncu: LSU 20 % + tensor-core
//   operand reads 29 % of the pipe; L2 -> SM 3.4 GB per 4736-tile launch against 6.45 GB).
// The feature segments (PE of the first / skip layer, direction + appearance of the view layer) stay SS-mode MMAs: their
// 16-column feature slices of both tiles ride in the same ring stage as the matching weight slice.
//
// Roles (608 threads): warps 0..15 epilogue (all 16 drain one accumulator half at a time, 32 columns per warp), warp 16 TMA
// producer, warps 17 / 18 MMA issuers of tile slot 0 / 1.  What paces the kernel once shared memory is out of the way is the
// INSTRUCTION CHAIN of the single-thread roles, measured step by step in round 2 (probe: 148 x 32 tiles, TFLOP/s):
//   719  one issuer, one table entry per ring stage and tile slot (~140 SASS instructions per 4 MMAs)
//   925  one entry per block (GEMM, N-half, segment, slot), per-stage inner loop
//  1014  epilogue specialised at compile time (was 3x more control than arithmetic instructions), TMEM loads pipelined
//  1039  the common blocks as ONE straight-line asm sequence (16 MMAs + 4 commits)
//  1097  32-bit descriptor arithmetic inside that sequence (a descriptor changes only in its low word): 240 -> 135 instructions
//  1168  one issuer per tile slot (each decodes, waits and builds operands while the other issues)
//  1197  issue token between the two issuers (strict X, Y alternation: free-running issuers phase-lock - both blocks
//        interleave in the pipe, finish together, and both slots wait for the epilogue at once)
//  1220  first-half epilogue: both TMEM loads back to back, `d_free` given before any arithmetic
// against 1020 for tc_mlp_pp_kernel on the same box.  scripts/probes/mma_chain_probe.cu: one thread issuing back-to-back reaches
// the 64 clk / MMA floor (N = 128), dependent accumulation chain or not, A from TMEM or shared memory;
// scripts/probes/tmem_ld_probe.cu: 16 warps drain tensor memory at ~690 B/clk/SM (a 64 KiB accumulator half in ~100 clk), and a
// continuous drain halves the MMA rate - neither is the limit here.  What is left (ncu, probe): epilogue warps busy ~70 % of
// the time (~150 instructions per warp and accumulator half), issuers waiting for the drain of their accumulator.
// Rejected on hardware: one group of 8 epilogue warps per tile slot (1085: the drain of a half takes longer per warp and
// both slots stall on it).
//
// The TMA producer and the issuers walk HOST-built tables (tp_build_program): one 16-byte entry per ring stage (producer) /
// per block (issuers).
#pragma once

constexpr int kTPStageBytes = 16384;
constexpr int kTPStageCols = 64;      // K columns of an activation stage (x <= 128 weight rows)
constexpr int kTPXCols = 16;          // K columns of a feature stage
constexpr int kTPX0Off = 8192;        // feature slice of tile slot 0 inside a feature stage (slot 1: + 4096)
constexpr int kTPXBytes = kTPXCols * kTileM * 2;
constexpr int kTPMaxStages = 12;
constexpr int kTPMaxProg = 320;       // issuer entries (table is padded by 2)
constexpr int kTPMaxLoads = 160;      // producer entries
constexpr int kTPThreads = kPPThreads + 32;   // 16 epilogue warps, TMA producer, two MMA issuers (one per tile slot)
enum { TF_FIRST = 2, TF_LAST = 4, TF_FROM_X = 8, TF_WAIT_A = 16 };

struct TPLayout {
    int ring, f32, f32_stride, sigp, bars, prog, loads, total, stages;
};

__host__ __device__ inline TPLayout tp_layout(const TcPlan& p) {
    TPLayout s;
    s.f32_stride = ((p.f32_floats * 4 + 15) / 16) * 16;
    const int fixed = s.f32_stride + 2048 + 512 + (kTPMaxProg + kTPMaxLoads) * 16;
    int st = (kSmemMax - fixed) / kTPStageBytes;
    if (st > kTPMaxStages) st = kTPMaxStages;
    s.stages = st;
    s.ring = 0;
    s.f32 = st * kTPStageBytes;
    s.sigp = s.f32 + s.f32_stride;
    s.bars = s.sigp + 2048;
    s.prog = s.bars + 512;
    s.loads = s.prog + kTPMaxProg * 16;
    s.total = s.loads + kTPMaxLoads * 16;
    return s;
}

// bytes of GEMM g's half-major image ([N-half][K/8][nw][8] fp16, nw = min(N, 128))
__host__ __device__ inline int tp_gemm_bytes(const TcGemm& g) {
    const int nw = g.n < 128 ? g.n : 128, nh = (g.n + 127) / 128;
    return (g.k[0] + (g.nseg > 1 ? g.k[1] : 0)) * nw * nh * 2;
}

// up to four K = 16 steps against one ring stage, A from TENSOR MEMORY (8 columns per step); one elected lane issues and,
// if `release`, frees the stage.
__device__ __forceinline__ void tp_stage_tmem(uint32_t d_tmem, uint32_t a_tmem, uint64_t bd, uint64_t b_step, uint32_t idesc,
                                              uint32_t accum, uint32_t nk, uint32_t empty_bar, uint32_t release) {
    asm volatile(
        "{\n\t.reg .pred e, p, q1, q2, q3, rl;\n\t.reg .b64 b1, b2, b3;\n\t.reg .b32 a1, a2, a3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.gt.and.u32 q1, %6, 1, e;\n\t"
        "setp.gt.and.u32 q2, %6, 2, e;\n\t"
        "setp.gt.and.u32 q3, %6, 3, e;\n\t"
        "setp.ne.and.b32 rl, %8, 0, e;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u64 b1, %2, %3;\n\tadd.u64 b2, b1, %3;\n\tadd.u64 b3, b2, %3;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %4, p;\n\t"
        "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %4, 1;\n\t"
        "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %4, 1;\n\t"
        "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %4, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bd), "l"(b_step), "r"(idesc), "r"(accum), "r"(nk), "r"(empty_bar), "r"(release)
        : "memory");
}
// one K = 16 step with A from shared memory (feature slice)
__device__ __forceinline__ void tp_stage_smem(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum,
                                              uint32_t empty_bar, uint32_t release) {
    asm volatile(
        "{\n\t.reg .pred e, p, rl;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.and.b32 rl, %6, 0, e;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum), "r"(empty_bar), "r"(release)
        : "memory");
}
// A whole activation block in ONE issue sequence: 4 ring stages x 4 K = 16 steps (16 MMAs, A from tensor memory) and, if
// `release`, the 4 commits that free the stages.  The issuer's instruction count per MMA - not the tensor pipe - paces this
// kernel (scripts/probes/mma_chain_probe.cu: one thread keeps the pipe at its 64 clk / MMA floor only if it spends well under 64
// clk per MMA), so the common block (K = 256) is straight-line code on 32-BIT operands: a shared-memory descriptor changes only
// in its low word (start address >> 4 | LBO << 16; the high word - SBO, version - is constant), b0..b3 = low words of the B
// descriptors of the four stages, b_step = 2 nw (one K = 16 step).
__device__ __forceinline__ void tp_block_tmem16(uint32_t d_tmem, uint32_t a_tmem, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3,
                                                uint32_t b_step, uint32_t b_hi, uint32_t idesc, uint32_t accum, uint32_t e0, uint32_t e1,
                                                uint32_t e2, uint32_t e3, uint32_t release) {
    asm volatile(
        "{\n\t.reg .pred e, p, rl;\n\t.reg .b64 q<16>;\n\t.reg .b32 a<16>, l<16>;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "setp.ne.and.b32 rl, %14, 0, e;\n\t"
        "add.u32 l1, %2, %6;\n\tadd.u32 l2, l1, %6;\n\tadd.u32 l3, l2, %6;\n\t"
        "add.u32 l5, %3, %6;\n\tadd.u32 l6, l5, %6;\n\tadd.u32 l7, l6, %6;\n\t"
        "add.u32 l9, %4, %6;\n\tadd.u32 l10, l9, %6;\n\tadd.u32 l11, l10, %6;\n\t"
        "add.u32 l13, %5, %6;\n\tadd.u32 l14, l13, %6;\n\tadd.u32 l15, l14, %6;\n\t"
        "mov.b64 q0, {%2, %7};\n\tmov.b64 q1, {l1, %7};\n\tmov.b64 q2, {l2, %7};\n\tmov.b64 q3, {l3, %7};\n\t"
        "mov.b64 q4, {%3, %7};\n\tmov.b64 q5, {l5, %7};\n\tmov.b64 q6, {l6, %7};\n\tmov.b64 q7, {l7, %7};\n\t"
        "mov.b64 q8, {%4, %7};\n\tmov.b64 q9, {l9, %7};\n\tmov.b64 q10, {l10, %7};\n\tmov.b64 q11, {l11, %7};\n\t"
        "mov.b64 q12, {%5, %7};\n\tmov.b64 q13, {l13, %7};\n\tmov.b64 q14, {l14, %7};\n\tmov.b64 q15, {l15, %7};\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\tadd.u32 a4, %1, 32;\n\tadd.u32 a5, %1, 40;\n\t"
        "add.u32 a6, %1, 48;\n\tadd.u32 a7, %1, 56;\n\tadd.u32 a8, %1, 64;\n\tadd.u32 a9, %1, 72;\n\tadd.u32 a10, %1, 80;\n\t"
        "add.u32 a11, %1, 88;\n\tadd.u32 a12, %1, 96;\n\tadd.u32 a13, %1, 104;\n\tadd.u32 a14, %1, 112;\n\tadd.u32 a15, %1, 120;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], q0, %8, p;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], q1, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], q2, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], q3, %8, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a4], q4, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a5], q5, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a6], q6, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a7], q7, %8, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a8], q8, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a9], q9, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a10], q10, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a11], q11, %8, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a12], q12, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a13], q13, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a14], q14, %8, 1;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], [a15], q15, %8, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(b_step), "r"(b_hi), "r"(idesc), "r"(accum), "r"(e0), "r"(e1),
          "r"(e2), "r"(e3), "r"(release)
        : "memory");
}
// A whole feature block: 5 ring stages x one K = 16 step, A (this tile slot's feature slice) and B from the same stage.
__device__ __forceinline__ void tp_block_smem5(uint32_t d_tmem, uint64_t ad0, uint64_t ad1, uint64_t ad2, uint64_t ad3, uint64_t ad4,
                                               uint64_t b_minus_a, uint32_t idesc, uint32_t accum, uint32_t e0, uint32_t e1, uint32_t e2,
                                               uint32_t e3, uint32_t e4, uint32_t release) {
    asm volatile(
        "{\n\t.reg .pred e, p, rl;\n\t.reg .b64 b<5>;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %8, 0;\n\t"
        "setp.ne.and.b32 rl, %14, 0, e;\n\t"
        "add.u64 b0, %1, %6;\n\tadd.u64 b1, %2, %6;\n\tadd.u64 b2, %3, %6;\n\tadd.u64 b3, %4, %6;\n\tadd.u64 b4, %5, %6;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %1, b0, %7, p;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %2, b1, %7, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %3, b2, %7, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %4, b3, %7, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %5, b4, %7, 1;\n\t"
        "@rl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t}"
        ::"r"(d_tmem), "l"(ad0), "l"(ad1), "l"(ad2), "l"(ad3), "l"(ad4), "l"(b_minus_a), "r"(idesc), "r"(accum), "r"(e0), "r"(e1), "r"(e2),
          "r"(e3), "r"(e4), "r"(release)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// tcgen05.wait::ld that also pins the 16 destination registers of an earlier tcgen05.ld behind it (register uses are otherwise
// free to move across an asm statement): lets a second load stay in flight under the arithmetic on the first.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    // no "memory" clobber: the wait orders tensor-memory -> register traffic only, so the bias loads of the next piece may be
    // scheduled above it
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 16-column piece of the epilogue for one accumulator row: + bias -> (ReLU) -> fp16 pairs; returns the partial sigma dot
// product if kSigma.  (ReLU after the fp16 rounding when no fp32 value is needed: max(round(x), 0) == round(max(x, 0)).)
template <bool kRelu, bool kSigma>
__device__ __forceinline__ float tp_piece16(const uint32_t (&v)[16], const float* __restrict__ bias16, const float* __restrict__ sw16,
                                            uint32_t* hp8) {
    float sacc = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 b = reinterpret_cast<const float4*>(bias16)[i];
        float f0 = __uint_as_float(v[4 * i + 0]) + b.x, f1 = __uint_as_float(v[4 * i + 1]) + b.y;
        float f2 = __uint_as_float(v[4 * i + 2]) + b.z, f3 = __uint_as_float(v[4 * i + 3]) + b.w;
        if (kSigma) {
            f0 = fmaxf(f0, 0.0f); f1 = fmaxf(f1, 0.0f); f2 = fmaxf(f2, 0.0f); f3 = fmaxf(f3, 0.0f);
            const float4 s4 = reinterpret_cast<const float4*>(sw16)[i];
            sacc = fmaf(f0, s4.x, sacc); sacc = fmaf(f1, s4.y, sacc);
            sacc = fmaf(f2, s4.z, sacc); sacc = fmaf(f3, s4.w, sacc);
        }
        uint32_t p0 = pack_h2(f0, f1), p1 = pack_h2(f2, f3);
        if (kRelu && !kSigma) {
            const __half2 z = __float2half2_rn(0.0f);
            __half2 h0 = __hmax2(*reinterpret_cast<__half2*>(&p0), z), h1 = __hmax2(*reinterpret_cast<__half2*>(&p1), z);
            p0 = *reinterpret_cast<const uint32_t*>(&h0);
            p1 = *reinterpret_cast<const uint32_t*>(&h1);
        }
        hp8[2 * i] = p0;
        hp8[2 * i + 1] = p1;
    }
    return sacc;
}

// Warps 0..15 epilogue, 16 TMA producer, 17 / 18 MMA issuers of tile slot 0 / 1.
__global__ void __launch_bounds__(kTPThreads, 1) tc_mlp_tp_kernel(const TcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const TPLayout SL = tp_layout(P);
    unsigned char* ring = smem + SL.ring;
    float* F32 = reinterpret_cast<float*>(smem + SL.f32);
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;             // [12]
    uint64_t* empty = bars + 12;       // [12]
    uint64_t* acc_full = bars + 24;    // [2] accumulator half of tile slot s complete
    uint64_t* d_free = bars + 26;      // [2] ... drained by the epilogue
    uint64_t* a_ready = bars + 28;     // [2] A operand of tile slot s written (next GEMM may start)
    uint64_t* f32_full = bars + 30;
    uint64_t* f32_empty = bars + 31;
    uint64_t* turn = bars + 32;        // [2] issue token of the two MMA issuers (strict alternation X, Y, X, Y ...)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);
    uint4* PROG = reinterpret_cast<uint4*>(smem + SL.prog);
    uint4* LOADS = reinterpret_cast<uint4*>(smem + SL.loads);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;
    const int n_prog = A.m.sigma_only ? A.tp_n[1] : A.tp_n[0];
    const int n_loads = A.m.sigma_only ? A.tp_n[3] : A.tp_n[2];
    const float* SW = F32 + P.sigma_w_off;
    const float* RGBB = F32 + P.g[P.n_gemm - 1].bias_off;

    for (int i = threadIdx.x; i < n_prog + 2; i += blockDim.x) PROG[i] = i < n_prog ? A.tp_prog[i] : make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < n_loads; i += blockDim.x) LOADS[i] = A.tp_prog[kTPMaxProg + i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTPMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }     // two issuers release a stage
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&d_free[i], kEpiWarps);
            mbar_init(&a_ready[i], kEpiWarps);
        }
        mbar_init(f32_full, 1);
        mbar_init(f32_empty, kEpiWarps);
        mbar_init(&turn[0], 1);
        mbar_init(&turn[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpProd) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
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
    // a CTA works on PAIRS of adjacent tiles (2p, 2p+1): buckets are 256-row aligned, so both belong to one sub-module
    const int64_t n_pairs = (n_tiles + 1) / 2;
    const uint32_t nst = (uint32_t)SL.stages;

    if (warp == kWarpProd) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, fph_e = 0, ahead = 0;
            int last_sub = -1;
            const uint32_t f32_bytes = (uint32_t)SL.f32_stride;
            const uint32_t empty_pa = smem_u32(empty), full_pa = smem_u32(full), ring_pa = smem_u32(ring);
            const int64_t xtile_bytes = (int64_t)(P.kpe + P.kaux) * kTileM * 2;
            for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
                const int64_t t0 = 2 * pr;
                const int sub0 = sub_of(t0);
                const unsigned char* wsub = A.tpack + (size_t)sub0 * A.tp_sub_bytes;
                const unsigned char* xt0 = reinterpret_cast<const unsigned char*>(A.ximg) + t0 * xtile_bytes;
                const bool valid1 = t0 + 1 < n_tiles;
                if (sub0 != last_sub) {
                    if (last_sub >= 0) { mbar_wait(f32_empty, fph_e); fph_e ^= 1; }
                    const unsigned char* fsrc = A.wpack + (size_t)sub0 * P.sub_bytes + (size_t)P.f32_off;
                    mbar_expect_tx(f32_full, f32_bytes);
                    bulk_g2s(reinterpret_cast<unsigned char*>(F32), fsrc, f32_bytes, f32_full);
                    last_sub = sub0;
                }
                const uint32_t xbytes = valid1 ? 2u * kTPXBytes : (uint32_t)kTPXBytes;
                for (int e = 0; e < n_loads; ++e) {
                    const uint4 E = LOADS[e];
                    const uint32_t cur = stage;
                    if (!ahead) mbar_wait_a(empty_pa + 8u * cur, phase ^ 1);
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                    ahead = mbar_test_a(empty_pa + 8u * stage, phase ^ 1);
                    const uint32_t bar = full_pa + 8u * cur, dst = ring_pa + cur * (uint32_t)kTPStageBytes;
                    const bool has_x = E.z != 0xFFFFFFFFu;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(E.y + (has_x ? xbytes : 0u)) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(wsub + E.x), "r"(E.y), "r"(bar) : "memory");
                    if (has_x) {
                        const unsigned char* xs = xt0 + ((int64_t)E.z << 4);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst + (uint32_t)kTPX0Off), "l"(xs), "r"((uint32_t)kTPXBytes), "r"(bar) : "memory");
                        if (valid1)
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst + (uint32_t)(kTPX0Off + kTPXBytes)), "l"(xs + xtile_bytes), "r"((uint32_t)kTPXBytes), "r"(bar)
                                         : "memory");
                    }
                }
            }
        }
    } else if (warp == kWarpMma || warp == kWarpMma + 1) {
        // =========================== MMA issuers: one warp per tile slot (one elected lane issues) ===========================
        // One table entry per BLOCK = the ring stages of one (GEMM, N-half, segment).  Both issuers walk the same table and the
        // same ring stages, each for its own tile slot (accumulator, A operand, barriers); a stage is free when BOTH have
        // committed it (`empty` counts 2).  An issuer's own instruction chain - not the tensor pipe, not the data - paces
        // this kernel (ncu: a single issuer never spins on a barrier, it is busy ~5 clk per instruction, ~250 instructions per
        // 16-MMA block): two issuers halve that, and the alternation X.h0, Y.h0, X.h1, Y.h1 follows from the data
        // dependencies (a slot's next block waits for its epilogue; the other slot's MMAs fill the tensor pipe meanwhile).
        const uint32_t sl = (uint32_t)(warp - kWarpMma);
        uint32_t stage = 0, phase = 0, ahead = 0, dph = 0, aph = 0;
        const uint32_t ring_base = smem_u32(ring);
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const uint32_t acc_full_a = smem_u32(acc_full) + 8u * sl, d_free_a = smem_u32(d_free) + 8u * sl, a_ready_a = smem_u32(a_ready) + 8u * sl;
        const uint64_t xd = make_desc(ring_base + (uint32_t)kTPX0Off + sl * (uint32_t)kTPXBytes, kTileM * 16, 128);
        const uint64_t bd_base = make_desc(ring_base, 0, 128);                 // LBO (= nw * 16 bytes) is added per entry
        const uint64_t st_step = (uint64_t)(kTPStageBytes >> 4);
        const uint32_t d_tmem = tmem_base + sl * 256u;
        // Issue token: the two issuers alternate block by block (X, Y, X, Y ...), i.e. the tensor pipe sees the order of the
        // one-issuer schedule - one slot's block runs while the other slot's accumulator is drained - but each issuer decodes
        // its entry, waits for its dependencies and builds its operands while the other one issues.  Free-running issuers
        // phase-lock instead: both blocks interleave in the pipe, finish together, and both slots wait for the epilogue at once.
        const uint32_t turn_mine = smem_u32(&turn[sl]), turn_other = smem_u32(&turn[sl ^ 1u]);
        uint32_t tph = sl ? 0u : 1u;       // issuer 0 starts (fresh barrier: the wait on parity 1 passes)
        for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
            const bool valid1 = 2 * pr + 1 < n_tiles;
            if (sl && !valid1) break;       // odd tile count: the last pair of the launch has no second tile
            uint4 E = PROG[0];
            for (int e = 0; e < n_prog; ++e) {
                const uint4 En = PROG[e + 1];
                const uint32_t fl = E.z >> 20;
                const uint32_t nw = E.z & 0xFFFu, ns = (E.z >> 12) & 0xFu, nk_last = (E.z >> 16) & 0xFu, idesc = E.y;
                uint32_t a_t = d_tmem + 128u + E.x;
                E = En;
                uint32_t accum = 1;
                if (fl & TF_FIRST) {
                    // the previous contents of this slot's accumulator have been read by the epilogue (fresh barrier: passes)
                    mbar_wait_a(d_free_a, dph ^ 1);
                    dph ^= 1;
                    accum = 0;
                }
                if (fl & TF_WAIT_A) {
                    // this GEMM's A operand (the previous GEMM's output) is in tensor memory
                    mbar_wait_a(a_ready_a, aph);
                    aph ^= 1;
                }
                const uint64_t bdn = bd_base + ((uint64_t)nw << 16);
                const uint64_t b_step = (uint64_t)(2u * nw);
                const uint32_t first = stage;
                if (ns == 4u && nk_last == 4u && !(fl & TF_FROM_X)) {
                    // ---- the common block (K = 256 activations): 16 MMAs in one straight-line issue sequence
                    const uint32_t c0 = stage, c1 = c0 + 1 == nst ? 0u : c0 + 1, c2 = c1 + 1 == nst ? 0u : c1 + 1, c3 = c2 + 1 == nst ? 0u : c2 + 1;
                    const uint32_t p0 = phase, p1 = c1 < c0 ? p0 ^ 1u : p0, p2 = c2 < c0 ? p0 ^ 1u : p0, p3 = c3 < c0 ? p0 ^ 1u : p0;
                    if (!ahead) mbar_wait_a(full_a + 8u * c0, p0);
                    mbar_wait_a(full_a + 8u * c1, p1);
                    mbar_wait_a(full_a + 8u * c2, p2);
                    mbar_wait_a(full_a + 8u * c3, p3);
                    const uint32_t nx = c3 + 1 == nst ? 0u : c3 + 1;
                    if (nx < stage) phase ^= 1;
                    stage = nx;
                    ahead = mbar_test_a(full_a + 8u * stage, phase);
                    tc_fence_after();
                    const uint32_t bl = (uint32_t)bdn, st32 = (uint32_t)st_step;
                    if (valid1) { mbar_wait_a(turn_mine, tph); tph ^= 1; }
                    tp_block_tmem16(d_tmem, a_t, bl + c0 * st32, bl + c1 * st32, bl + c2 * st32, bl + c3 * st32, 2u * nw, (uint32_t)(bdn >> 32), idesc,
                                    accum, empty_a + 8u * c0, empty_a + 8u * c1, empty_a + 8u * c2, empty_a + 8u * c3, 1u);
                } else if (ns == 5u && (fl & TF_FROM_X)) {
                    // ---- the feature block (80 encoding columns): 5 MMAs, A and B of a step in the same stage
                    uint32_t c[5], ph[5];
                    c[0] = stage; ph[0] = phase;
#pragma unroll
                    for (int i = 1; i < 5; ++i) { c[i] = c[i - 1] + 1 == nst ? 0u : c[i - 1] + 1; ph[i] = c[i] < c[0] ? phase ^ 1u : phase; }
                    if (!ahead) mbar_wait_a(full_a + 8u * c[0], ph[0]);
#pragma unroll
                    for (int i = 1; i < 5; ++i) mbar_wait_a(full_a + 8u * c[i], ph[i]);
                    const uint32_t nx = c[4] + 1 == nst ? 0u : c[4] + 1;
                    if (nx < stage) phase ^= 1;
                    stage = nx;
                    ahead = mbar_test_a(full_a + 8u * stage, phase);
                    tc_fence_after();
                    if (valid1) { mbar_wait_a(turn_mine, tph); tph ^= 1; }
                    tp_block_smem5(d_tmem, xd + (uint64_t)c[0] * st_step, xd + (uint64_t)c[1] * st_step, xd + (uint64_t)c[2] * st_step,
                                   xd + (uint64_t)c[3] * st_step, xd + (uint64_t)c[4] * st_step, bdn - xd, idesc, accum, empty_a + 8u * c[0],
                                   empty_a + 8u * c[1], empty_a + 8u * c[2], empty_a + 8u * c[3], empty_a + 8u * c[4], 1u);
                } else {
                    // ---- any other block shape (narrower networks, the colour head): one stage per iteration
                    if (valid1) { mbar_wait_a(turn_mine, tph); tph ^= 1; }
                    for (uint32_t s2 = 0; s2 < ns; ++s2) {
                        const uint32_t cur = stage;
                        if (!ahead) mbar_wait_a(full_a + 8u * cur, phase);
                        if (++stage == nst) { stage = 0; phase ^= 1; }
                        ahead = mbar_test_a(full_a + 8u * stage, phase);
                        tc_fence_after();
                        const uint64_t so = (uint64_t)cur * st_step;
                        if (fl & TF_FROM_X) tp_stage_smem(d_tmem, xd + so, bdn + so, idesc, accum, empty_a + 8u * cur, 1u);
                        else {
                            tp_stage_tmem(d_tmem, a_t, bdn + so, b_step, idesc, accum, s2 + 1 == ns ? nk_last : 4u, empty_a + 8u * cur, 1u);
                            a_t += 32u;
                        }
                        accum = 1;
                    }
                }
                if (valid1) {
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(turn_other) : "memory");
                    __syncwarp();
                } else {
                    // no second tile: this issuer also gives the other issuer's release of the block's stages
                    uint32_t c = first;
                    for (uint32_t s2 = 0; s2 < ns; ++s2) {
                        commit_elect(empty_a + 8u * c);
                        if (++c == nst) c = 0;
                    }
                }
                if (fl & TF_LAST) commit_elect(acc_full_a);
            }
        }
    } else {
        // =========================== epilogue (16 warps) ===========================
        const int q = warp & 3;          // TMEM lane quarter
        const int part = warp >> 2;      // 32-column piece of the 128-column accumulator half
        const int r = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t aphm = 0, fph0 = 0;      // aphm: parity of acc_full[s] in bit s
        const uint32_t acc_full_a = smem_u32(acc_full);
        int last_sub = -1;
        const int L = P.L;
        for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
            const int64_t t0 = 2 * pr;
            const bool valid1 = t0 + 1 < n_tiles;
            {
                const int sub0 = sub_of(t0);
                if (sub0 != last_sub) {
                    if (last_sub >= 0) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(f32_empty);
                    }
                    mbar_wait(f32_full, fph0);
                    fph0 ^= 1;
                    last_sub = sub0;
                }
            }
            float sigma_[2] = {0.0f, 0.0f}, sacc_[2] = {0.0f, 0.0f};
            uint32_t keep0[16] = {}, keep1[16] = {};      // packed h0 half of the current layer's output, per tile slot
            for (int gi = 0; gi < n_gemm; ++gi) {
                // per-GEMM scalars in registers (the plan lives in the kernel parameters: dynamic constant-bank reads)
                const int g_n = P.g[gi].n, g_epi = P.g[gi].epi;
                const int nh = (g_n + 127) >> 7;
                const float* bias = F32 + P.g[gi].bias_off;
                const bool want_sigma = g_epi == EPI_RELU_SIGMA;
                const bool publish = g_epi != EPI_RGB && !(want_sigma && A.m.sigma_only);
                if (g_epi == EPI_RGB) {
                    // colour head: one code path for both tile slots (nothing is stashed here)
                    for (int sl = 0; sl < (valid1 ? 2 : 1); ++sl) {
                        mbar_wait(&acc_full[sl], (aphm >> sl) & 1u);
                        aphm ^= 1u << sl;
                        tc_fence_after();
                        const int64_t slot = (t0 + sl) * kTileM + r;
                        const int64_t row = slot < n_slots ? (A.m.slot_row ? (int64_t)A.m.slot_row[slot] : slot) : -1;
                        uint32_t v[32];
                        tmem_ld32(t_lane + (uint32_t)sl * 256u, v);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_free[sl]);
                        if (part == 0 && row >= 0)
                            tc_emit_rgb(A.m, A.m.nd.affine ? sub_of(t0 + sl) : 0, row, slot, v, RGBB, sl ? sigma_[1] : sigma_[0], nullptr);
                    }
                    continue;
                }
                for (int h = 0; h < nh; ++h) {
                    const int n0 = 128 * h + 32 * part;      // first output channel of this warp's piece
                    const bool active = n0 < g_n;
                    const bool last_h = h == nh - 1;
                    const int nb = active ? n0 : 0;          // inactive warps (N < 128 halves) run on don't-care columns, no stores
                    const float* b16 = bias + nb;
                    const float* s16 = SW + nb;
                    const uint32_t c_src = (uint32_t)(active ? 32 * part : 0);
                    const bool stash_out = publish && nh == 2;
                    // one tile slot's share of this accumulator half.  Generic lambda: the tile slot and the epilogue flavour are
                    // compile-time constants - the stash arrays stay in registers and the arithmetic is branch-free.
                    auto do_slot = [&](auto slc, auto reluc, auto sigc, auto lastc, uint32_t (&keep)[16]) {
                        constexpr int sl = decltype(slc)::value;
                        constexpr bool kRelu = decltype(reluc)::value, kSigma = decltype(sigc)::value, kLast = decltype(lastc)::value;
                        mbar_wait_a(acc_full_a + 8u * sl, (aphm >> sl) & 1u);
                        aphm ^= 1u << sl;
                        tc_fence_after();
                        const uint32_t t_acc = t_lane + (uint32_t)sl * 256u;
                        uint32_t v0[16], v1[16];
                        tmem_ld16(t_acc + c_src, v0);
                        // last half: every MMA that read this slot's A operand has completed - the stashed h0 half goes out now,
                        // under the loads, instead of at the tail of the dependent chain
                        if (kLast && stash_out) tmem_st16(t_acc + 128u + (uint32_t)(16 * part), keep);
                        if (kLast) {
                            tmem_ld_wait16(v0);
                            tmem_ld16(t_acc + c_src + 16u, v1);      // in flight under the arithmetic on v0
                        } else {
                            // first half: the issuer of this slot waits for the DRAIN of the accumulator (its next block is the
                            // layer's second half) - both loads go out back to back and `d_free` is given before any arithmetic
                            tmem_ld16(t_acc + c_src + 16u, v1);
                            tmem_ld_wait16(v0);
                            tmem_ld_wait16(v1);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&d_free[sl]);
                        }
                        if (kLast) {
                            uint32_t hp[16];
                            float sacc = tp_piece16<kRelu, kSigma>(v0, b16, s16, hp);
                            tmem_ld_wait16(v1);
                            tc_fence_before();                   // this warp no longer needs the accumulator
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&d_free[sl]);
                            sacc += tp_piece16<kRelu, kSigma>(v1, b16 + 16, s16 + 16, hp + 8);
                            if (publish) {
                                if (active) tmem_st16(t_acc + 128u + (uint32_t)(n0 >> 1), hp);
                                tmem_st_wait();
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&a_ready[sl]);
                            }
                            if (kSigma) {
                                SIGP[part * kTileM + r] = sacc_[sl] + (active ? sacc : 0.0f);
                                sacc_[sl] = 0.0f;
                                asm volatile("bar.sync 1, 512;" ::: "memory");
                                if (part == 0) {
                                    const int64_t slot = (t0 + sl) * kTileM + r;
                                    const int64_t row = slot < n_slots ? (A.m.slot_row ? (int64_t)A.m.slot_row[slot] : slot) : -1;
                                    float sv = ((SIGP[r] + SIGP[kTileM + r]) + (SIGP[2 * kTileM + r] + SIGP[3 * kTileM + r])) + SW[L];
                                    if (A.m.sigma_noise && row >= 0) sv = sv + A.m.sigma_noise[row];
                                    const float sg = A.m.nd.softplus ? mn_softplus_shifted(sv) : fmaxf(sv, 0.0f);
                                    sigma_[sl] = sg;
                                    if (A.m.sigma_only && row >= 0) {
                                        const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                                        A.m.out[o] = A.m.slot_w ? sg * A.m.slot_w[slot] : sg;
                                    }
                                }
                                asm volatile("bar.sync 1, 512;" ::: "memory");   // SIGP is reused by the other tile slot
                            }
                        } else {
                            // first half of a two-half layer: straight into the stash (all four column pieces are active)
                            float sacc = tp_piece16<kRelu, kSigma>(v0, b16, s16, keep);
                            sacc += tp_piece16<kRelu, kSigma>(v1, b16 + 16, s16 + 16, keep + 8);
                            if (kSigma) sacc_[sl] += sacc;
                        }
                    };
                    using T = std::true_type;
                    using F = std::false_type;
                    using S0 = std::integral_constant<int, 0>;
                    using S1 = std::integral_constant<int, 1>;
                    if (g_epi == EPI_RELU) {
                        if (last_h) { do_slot(S0{}, T{}, F{}, T{}, keep0); if (valid1) do_slot(S1{}, T{}, F{}, T{}, keep1); }
                        else        { do_slot(S0{}, T{}, F{}, F{}, keep0); if (valid1) do_slot(S1{}, T{}, F{}, F{}, keep1); }
                    } else if (g_epi == EPI_LINEAR) {
                        if (last_h) { do_slot(S0{}, F{}, F{}, T{}, keep0); if (valid1) do_slot(S1{}, F{}, F{}, T{}, keep1); }
                        else        { do_slot(S0{}, F{}, F{}, F{}, keep0); if (valid1) do_slot(S1{}, F{}, F{}, F{}, keep1); }
                    } else {
                        if (last_h) { do_slot(S0{}, T{}, T{}, T{}, keep0); if (valid1) do_slot(S1{}, T{}, T{}, T{}, keep1); }
                        else        { do_slot(S0{}, T{}, T{}, F{}, keep0); if (valid1) do_slot(S1{}, T{}, T{}, F{}, keep1); }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    clk_stamp(A.desc_swap, 1);
    if (warp == kWarpProd) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// Host: the two tables of tc_mlp_tp_kernel for plan P.  prog[kTPMaxProg] issuer entries then loads[kTPMaxLoads] producer
// entries; counts = {issuer entries (all GEMMs), issuer entries (trunk only: sigma_only calls), producer (all), producer (trunk)}.
//   issuer   x = A column offset in tensor memory (0)   y = instruction descriptor
//            z = nw | stages << 12 | K-steps of the last stage << 16 | TF_* flags << 20   w = 2 * GEMM + N-half
//   producer x = byte offset inside the sub-module's image        y = weight bytes        z = feature offset / 16 or ~0
static bool tp_build_program(const TcPlan& P, std::vector<uint4>* table, int counts[4]) {
    std::vector<uint4> prog, loads;
    int woff = 0;
    counts[1] = counts[3] = -1;
    for (int gi = 0; gi < P.n_gemm; ++gi) {
        if (gi == P.n_trunk) { counts[1] = (int)prog.size(); counts[3] = (int)loads.size(); }
        const TcGemm& g = P.g[gi];
        const int nw = g.n < 128 ? g.n : 128, nh = (g.n + 127) / 128;
        const int K = g.k[0] + (g.nseg > 1 ? g.k[1] : 0);
        for (int h = 0; h < nh; ++h) {
            const int himg = woff + h * K * nw * 2;
            int kbase = 0;
            bool waited = false;
            for (int sgi = 0; sgi < g.nseg; ++sgi) {
                const bool fx = g.src[sgi] != SRC_H;
                const int step = fx ? kTPXCols : kTPStageCols;
                const int kk = g.k[sgi];
                const int ns = (kk + step - 1) / step;
                if (ns > kTPMaxStages - 4) return false;      // a block must leave ring stages for the next block's prefetch
                for (int j = 0; j < ns; ++j) {
                    const int k0 = j * step, kc = kk - k0 < step ? kk - k0 : step;
                    const unsigned xo = fx ? (unsigned)(((g.src[sgi] == SRC_XAUX ? P.kpe * kTileM * 2 : 0) + k0 * kTileM * 2) >> 4) : 0xFFFFFFFFu;
                    loads.push_back(make_uint4((unsigned)(himg + (kbase + k0) * nw * 2), (unsigned)(kc * nw * 2), xo, 0u));
                }
                // one issuer entry per block (both issuers walk the same table)
                const int kc_last = kk - (ns - 1) * step;
                {
                    unsigned fl = (sgi == 0 ? TF_FIRST : 0) | (sgi == g.nseg - 1 ? TF_LAST : 0) | (fx ? TF_FROM_X : 0);
                    if (!fx && h == 0) fl |= TF_WAIT_A;
                    // kind::f16, D = f32, K-major A and B, N >> 3 at [17,23), M >> 4 at [24,29)  (make_idesc)
                    const unsigned idesc = (1u << 4) | ((unsigned)(nw >> 3) << 17) | ((unsigned)(kTileM >> 4) << 24);
                    prog.push_back(make_uint4(0u, idesc, (unsigned)nw | ((unsigned)ns << 12) | ((unsigned)(kc_last >> 4) << 16) | (fl << 20),
                                              (unsigned)(gi * 2 + h)));
                }
                kbase += kk;
            }
            (void)waited;
        }
        woff += tp_gemm_bytes(g);
    }
    if (counts[1] < 0) { counts[1] = (int)prog.size(); counts[3] = (int)loads.size(); }
    counts[0] = (int)prog.size();
    counts[2] = (int)loads.size();
    if (counts[0] + 2 > kTPMaxProg || counts[2] > kTPMaxLoads) return false;
    table->assign(kTPMaxProg + kTPMaxLoads, make_uint4(0u, 0u, 0u, 0u));
    for (size_t i = 0; i < prog.size(); ++i) (*table)[i] = prog[i];
    for (size_t i = 0; i < loads.size(); ++i) (*table)[kTPMaxProg + i] = loads[i];
    return true;
}

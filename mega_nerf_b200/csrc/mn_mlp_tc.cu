// MN_PREC_TC_F16 / MN_PREC_TC_F16X3: the NeRF MLP (models/nerf.py:115-160) on the 5th-gen tensor cores.
//
//   tc_encode_kernel   sample rows -> fp16 feature tiles (positional encodings of xyz / dir, appearance
//                      embedding) written in the exact shared-memory operand image, one 128-row tile
//                      per CTA, coalesced 16-byte stores.
//   tc_mlp_kernel      persistent, warp-specialised: warp 0 = TMA producer (cp.async.bulk of weight
//                      K-slabs through an mbarrier ring and of the feature tiles), warp 1 = single-thread
//                      tcgen05.mma issuer (accumulators in TMEM), warps 2-5 = epilogue (tcgen05.ld ->
//                      bias/ReLU -> fp16 -> next layer's A operand in shared memory; heads -> HBM).
//                      Activations never leave the SM between layers.
//
// Operand layout (both A tiles and packed weights): K-major, no swizzle, "interleaved" core matrices:
//   element (row r, col k) of an R-row operand lives at byte (k/8)*(R*16) + r*16 + (k%8)*2,
// i.e. [K/8][R][8] fp16.  UMMA descriptor: LBO = R*16 (next 8-column chunk), SBO = 128 (next 8-row group).
// The epilogue's per-row 16-byte stores and the packer's images are contiguous in this layout, any K that
// is a multiple of 16 works, and no TMA tensor map is needed (plain 1-D bulk copies).
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "mn_model.cuh"

namespace {

constexpr int kTileM = 128;

// Optional in-kernel timeline (MN_TC_TRACE=1): CTA 0 records (event, tile-slot, gemm, clock) tuples.
__device__ unsigned long long g_trace[4 * 4096];
__device__ unsigned int g_trace_n[2];
__device__ __forceinline__ void trace_ev(int on, int who, int ev, int sl, int gi) {
    if (!(on & 1) || blockIdx.x != 0) return;
    const unsigned int i = atomicAdd(&g_trace_n[who], 1u);
    if (i < 2048) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[(who * 2048 + i) * 2 + 0] = ((unsigned long long)ev << 32) | ((unsigned long long)sl << 16) | (unsigned long long)gi;
        g_trace[(who * 2048 + i) * 2 + 1] = t;
    }
}
// SM clock during the kernel (MN_TC_TRACE=1): thread 0 of CTA 0 stamps (clock64, globaltimer) at kernel start and end.
__device__ unsigned long long g_clk[4];
__device__ __forceinline__ void clk_stamp(int on, int which) {
    if (!(on & 1) || blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_clk[2 * which] = (unsigned long long)clock64();
    g_clk[2 * which + 1] = t;
}
constexpr int kMaxStages = 4;
// weight ring geometry: single pass = 3 stages x 64 K-columns (32 KiB); split mode (H holds hi+lo planes) = 4 x 32 columns
__host__ __device__ constexpr int ring_slab_cols(bool split) { return split ? 32 : 64; }
__host__ __device__ constexpr int ring_stage_bytes(bool split) { return ring_slab_cols(split) * 256 * 2; }
constexpr int kMaxGemm = 16;
constexpr int kThreads = 576;   // producer warp, MMA warp, 16 epilogue warps
constexpr int kEpiWarps = 16;
// Warps 0..15 = epilogue (TMEM lane quarter = warp % 4), 16 = TMA producer, 17 = MMA issuer.  The scheduler favours the
// highest warp id on an SM sub-partition, so the latency-critical single-thread roles get the top ids.
constexpr int kWarpProd = 16;
constexpr int kWarpMma = 17;
constexpr int kPPThreads = kThreads;

enum { SRC_H = 0, SRC_XPE = 1, SRC_XAUX = 2 };
enum { EPI_RELU = 0, EPI_RELU_SIGMA = 1, EPI_LINEAR = 2, EPI_RGB = 3,
       // data-gradient chain (training, mn_train_tc.cuh): plain copy, ReLU mask from the activation tape, mask + sigma-head term
       EPI_D_LINEAR = 4, EPI_D_MASK = 5, EPI_D_MASK_SIGMA = 6 };
// kernel modes of tc_mlp_pp_kernel
enum { PP_INFER = 0, PP_TRAIN_FWD = 1, PP_DGRAD = 2 };

struct TcGemm {
    int n;           // MMA N
    int nseg;
    int src[2];
    int k[2];        // padded K columns per segment (multiple of 16)
    int w_off;       // byte offset of the weight image inside one precision plane of a sub-module
    int bias_off;    // float offset inside the sub-module's fp32 block
    int epi;
};

struct TcPlan {
    int n_gemm, n_trunk;
    TcGemm g[kMaxGemm];
    int kpe, kaux;         // padded feature-tile widths
    int plane_bytes;       // bytes of all weight images of one sub-module (one precision plane)
    int f32_floats;        // fp32 block: biases per GEMM (256 each) + sigma_w[L] + sigma_b
    int sigma_w_off;       // float offset of sigma_w in the fp32 block
    int sub_bytes;         // total bytes per sub-module: planes (hi[,lo]) + fp32 block
    int x_tile_bytes;      // bytes of one feature tile image (one plane)
    int L;
    int bstride;           // floats reserved per GEMM bias in the fp32 block (256; 512 for the 512-wide network)
    int f32_off;           // byte offset of the fp32 block inside one sub-module's pack (after the hi and lo planes)
};

int pad16(int x) { return (x + 15) / 16 * 16; }

bool build_plan(const NetDims& nd, TcPlan* p) {
    if (nd.L % 64 != 0 || (nd.L > 256 && nd.L != 512) || nd.L < 64 || nd.rgb_dim > 32 || nd.layers > 12) return false;
    if (nd.affine && nd.rgb_dim != 3) return false;
    TcPlan& P = *p;
    P = TcPlan{};
    P.L = nd.L;
    P.bstride = nd.L > 256 ? 512 : 256;
    P.kpe = pad16(nd.in_xyz);
    P.kaux = nd.aux > 0 ? pad16(nd.aux) : 0;
    int woff = 0, ng = 0;
    auto add = [&](int n, int s0, int k0, int s1, int k1, int epi) {
        TcGemm& g = P.g[ng];
        g.n = n;
        g.nseg = k1 > 0 ? 2 : 1;
        g.src[0] = s0; g.k[0] = k0; g.src[1] = s1; g.k[1] = k1;
        g.w_off = woff;
        g.bias_off = ng * P.bstride;
        g.epi = epi;
        woff += (k0 + k1) * n * 2;
        ++ng;
    };
    for (int i = 0; i < nd.layers; ++i) {
        const int epi = (i == nd.layers - 1) ? EPI_RELU_SIGMA : EPI_RELU;
        if (i == 0) add(nd.L, SRC_XPE, P.kpe, 0, 0, epi);
        else if ((nd.skip_mask >> i) & 1) add(nd.L, SRC_XPE, P.kpe, SRC_H, nd.L, epi);
        else add(nd.L, SRC_H, nd.L, 0, 0, epi);
    }
    P.n_trunk = ng;
    if (nd.has_dir_a) {
        add(nd.L, SRC_H, nd.L, 0, 0, EPI_LINEAR);
        add(nd.L / 2, SRC_H, nd.L, SRC_XAUX, P.kaux, EPI_RELU);
        add(32, SRC_H, nd.L / 2, 0, 0, EPI_RGB);
    } else {
        add(32, SRC_H, nd.L, 0, 0, EPI_RGB);
    }
    P.n_gemm = ng;
    P.plane_bytes = woff;
    P.sigma_w_off = ng * P.bstride;
    P.f32_floats = ng * P.bstride + nd.L + 4;
    P.f32_off = woff * 2;
    P.x_tile_bytes = (P.kpe + P.kaux) * kTileM * 2;
    return true;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            printf("mn_mlp_tc: mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
    // kind::f16: D=f32 (bit 4), A=B=f16 (0), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// One 16-column piece of the epilogue for one accumulator row: TMEM -> +bias -> (ReLU) -> fp16 (hi [, lo]) ->
// two 16-byte stores into the next layer's A operand.  Returns the partial sigma dot product if kSigma.
template <bool kSplit, bool kRelu, bool kSigma>
__device__ __forceinline__ float epi_piece16(uint32_t taddr, const float* __restrict__ bias16, const float* __restrict__ sw16,
                                             unsigned char* dst, size_t lo_off, bool store, unsigned char* gdst = nullptr) {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    const float4* b4 = reinterpret_cast<const float4*>(bias16);
    const float4 b0 = b4[0], b1 = b4[1], b2 = b4[2], b3 = b4[3];
    const float b[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
    tmem_ld_wait();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + b[i];
    float sacc = 0.0f;
    if (kSigma || kSplit) {
        if (kRelu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
        }
    }
    if (kSigma) {
        const float4* s4 = reinterpret_cast<const float4*>(sw16);
        const float4 s0 = s4[0], s1 = s4[1], s2 = s4[2], s3 = s4[3];
        const float sw[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc = fmaf(f[i], sw[i], sacc);
    }
    if (store) {
        uint32_t hi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) hi[e] = pack_h2(f[2 * e], f[2 * e + 1]);
        if (kRelu && !(kSigma || kSplit)) {
            // ReLU after the fp16 rounding (max(round(x),0) == round(max(x,0))): one HMNMX2 per pair
            const __half2 z = __float2half2_rn(0.0f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                __half2 h = __hmax2(*reinterpret_cast<__half2*>(&hi[e]), z);
                hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + kTileM * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        if (gdst) {     // training forward: the same two 16-byte pieces go to the activation tape (same image layout)
            *reinterpret_cast<uint4*>(gdst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(gdst + kTileM * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        }
        if (kSplit) {
            uint32_t lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi[e]));
                lo[e] = pack_h2(f[2 * e] - back.x, f[2 * e + 1] - back.y);
            }
            *reinterpret_cast<uint4*>(dst + lo_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(dst + lo_off + kTileM * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
    }
    return sacc;
}

// rgb head epilogue shared by the tensor-core kernels (nerf.py:152-160): bias, optional per-image affine appearance
// transform (3x4 matrix = affine(embedding_a[idx]), nerf.py:156-158), sigmoid when rgb_dim == 3, blend weight.
// v = the row's raw fp32 accumulators of the rgb GEMM.
__device__ __forceinline__ void tc_emit_rgb(const MlpArgs& m, int sub, int64_t row, int64_t slot, const uint32_t* v,
                                            const float* bias, float sigma, float* tape_rgb = nullptr) {
    const NetDims& nd = m.nd;
    const int64_t o = (m.scatter ? row : slot) * m.out_cols;
    const float w = m.slot_w ? m.slot_w[slot] : 1.0f;
    if (nd.affine && nd.app > 0) {
        const float* Pk = m.packed + (size_t)sub * m.lay.total;
        const float* emb = Pk + m.lay.emb;
        const float* aw = Pk + m.lay.aff_w;   // [app][12]
        int id = (int)m.src.index(row);
        id = min(max(id, 0), nd.app_count - 1);
        float T[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) T[q] = Pk[m.lay.aff_b + q];
        for (int j = 0; j < nd.app; ++j) {
            const float e = emb[(size_t)id * nd.app + j];
#pragma unroll
            for (int q = 0; q < 12; ++q) T[q] = fmaf(e, aw[j * 12 + q], T[q]);
        }
        const float r0 = __uint_as_float(v[0]) + bias[0], r1 = __uint_as_float(v[1]) + bias[1], r2 = __uint_as_float(v[2]) + bias[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = mn_sigmoid(fmaf(T[c * 4 + 2], r2, fmaf(T[c * 4 + 1], r1, T[c * 4 + 0] * r0)) + T[c * 4 + 3]);
            m.out[o + c] = m.slot_w ? x * w : x;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (c < nd.rgb_dim) {
                float x = __uint_as_float(v[c]) + bias[c];
                if (nd.rgb_dim == 3) x = mn_sigmoid(x);
                if (tape_rgb && c < 3) tape_rgb[c * kTileM] = x;      // training forward: colour before the blend weight
                m.out[o + c] = m.slot_w ? x * w : x;
            }
        }
    }
    m.out[o + nd.rgb_dim] = m.slot_w ? sigma * w : sigma;
}

// ------------------------------------------------------------------------------------------------
// weight packing: nn.Linear weight [N_src][K_src] fp32 -> image [K/8][N][8] fp16 (hi) and the residual (lo)
// ------------------------------------------------------------------------------------------------

// half-major image for the TS kernel: [N-half][K/8][nw][8] fp16 (nw = min(N,128))


// ------------------------------------------------------------------------------------------------
// feature tiles
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileM) tc_encode_kernel(const MlpArgs a, int kpe, int kaux, int split,
                                                           __half* __restrict__ ximg, int64_t plane_stride_halves) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __half* img = reinterpret_cast<__half*>(sm_raw);                  // hi image, then lo image
    const int ktot = kpe + kaux;
    const NetDims& nd = a.nd;
    const int t = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t slot0 = tile * kTileM;
    const int64_t n_slots = a.counters ? a.counters[CNT_NSLOTS] : a.B;
    if (slot0 >= n_slots) return;
    const int64_t slot = slot0 + t;
    int64_t row = -1;
    if (slot < n_slots) row = a.slot_row ? (int64_t)a.slot_row[slot] : slot;
    __half* lo_img = img + (size_t)ktot * kTileM;
    auto put = [&](int col, float v) {
        const int o = (col >> 3) * (kTileM * 8) + t * 8 + (col & 7);
        const __half h = __float2half_rn(v);
        img[o] = h;
        if (split) lo_img[o] = __float2half_rn(v - __half2float(h));
    };
    int sub = a.fixed_sub;
    if (a.counters) {
        sub = 0;
        while (sub + 1 < a.n_sub && slot0 >= a.counters[CNT_START + sub + 1]) ++sub;
    }
    if (row < 0) {
        for (int c = 0; c < ktot; ++c) put(c, 0.0f);
    } else {
        float x[4];
        double xp[4];
        for (int j = 0; j < nd.xyz_dim; ++j) { x[j] = a.src.xyz(row, j); put(j, x[j]); xp[j] = mn_pe_prescale(x[j]); }
        for (int k = 0; k < nd.nf_xyz; ++k)
            for (int j = 0; j < nd.xyz_dim; ++j) {
                float s, c;
                mn_pe_sincos(x[j], k, &s, &c);
                const int base = nd.xyz_dim + k * 2 * nd.xyz_dim;
                put(base + j, s);
                put(base + nd.xyz_dim + j, c);
            }
        for (int c = nd.in_xyz; c < kpe; ++c) put(c, 0.0f);
        if (kaux > 0) {
            int col = kpe;
            if (!a.sigma_only) {
                if (nd.nf_dir > 0) {
                    float d[3];
                    double dp[3];
                    for (int j = 0; j < 3; ++j) { d[j] = a.src.dir(row, j); put(col + j, d[j]); dp[j] = mn_pe_prescale(d[j]); }
                    for (int k = 0; k < nd.nf_dir; ++k)
                        for (int j = 0; j < 3; ++j) {
                            float s, c;
                            mn_pe_sincos(d[j], k, &s, &c);
                            put(col + 3 + k * 6 + j, s);
                            put(col + 3 + k * 6 + 3 + j, c);
                        }
                    col += nd.in_dir;
                }
                if (nd.app_in_dira) {
                    const float* emb = a.packed + (size_t)sub * a.lay.total + a.lay.emb;
                    int id = (int)a.src.index(row);
                    id = min(max(id, 0), nd.app_count - 1);
                    for (int j = 0; j < nd.app; ++j) put(col + j, emb[(size_t)id * nd.app + j]);
                    col += nd.app;
                }
            }
            for (int c = col; c < ktot; ++c) put(c, 0.0f);
        }
    }
    __syncthreads();
    const int nvec = ktot * kTileM * 2 / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(img);
    uint4* d4 = reinterpret_cast<uint4*>(ximg + tile * (int64_t)ktot * kTileM);
    for (int i = t; i < nvec; i += kTileM) d4[i] = s4[i];
    if (split) {
        const uint4* s4l = reinterpret_cast<const uint4*>(lo_img);
        uint4* d4l = reinterpret_cast<uint4*>(ximg + plane_stride_halves + tile * (int64_t)ktot * kTileM);
        for (int i = t; i < nvec; i += kTileM) d4l[i] = s4l[i];
    }
}


// sin / cos of 2^k x for consecutive k (nerf.py:19-25): every fourth band is evaluated with the accurate sincosf, the
// three bands after it by the double-angle identities.  Each doubling at most doubles the absolute error, so the
// result stays within ~8 fp32 ulps (5e-7) of the directly evaluated value - three orders of magnitude below the fp16
// rounding (2.4e-4) these features undergo on their way into the tensor-core operand tile.  (s, c) carry band k-1 in.
__device__ __forceinline__ void pe_band(float x, int k, float* s, float* c) {
    if ((k & 3) == 0) {
        mn_pe_sincos(x, k, s, c);
    } else {
        const float s0 = *s, c0 = *c;
        *s = 2.0f * s0 * c0;
        *c = 1.0f - 2.0f * s0 * s0;
    }
}

// Specialised encoder for the common network shape (compile-time channel counts): every thread builds its row's
// channels in registers and writes the tile image straight to global memory with 16-byte stores (thread t of a
// chunk writes bytes [t*16, t*16+16) -> fully coalesced); no shared-memory staging, no 2-byte bank-conflicted stores.
template <int XD, int NFX, int NFD, int APP>
__global__ void __launch_bounds__(kTileM) tc_encode_fast_kernel(const MlpArgs a, __half* __restrict__ ximg) {
    constexpr int IN_XYZ = XD * (1 + 2 * NFX);
    constexpr int KPE = (IN_XYZ + 15) / 16 * 16;
    constexpr int IN_DIR = NFD > 0 ? 3 + 6 * NFD : 0;
    constexpr int KAUX = (IN_DIR + APP + 15) / 16 * 16;
    const int t = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t slot0 = tile * kTileM;
    const int64_t n_slots = a.counters ? a.counters[CNT_NSLOTS] : a.B;
    if (slot0 >= n_slots) return;
    const int64_t slot = slot0 + t;
    int64_t row = -1;
    if (slot < n_slots) row = a.slot_row ? (int64_t)a.slot_row[slot] : slot;
    uint4* out = reinterpret_cast<uint4*>(ximg + tile * (int64_t)(KPE + KAUX) * kTileM) + t;   // + chunk * kTileM
    {
        float v[KPE];
#pragma unroll
        for (int c = 0; c < KPE; ++c) v[c] = 0.0f;
        if (row >= 0) {
#pragma unroll
            for (int j = 0; j < XD; ++j) {
                const float x = a.src.xyz(row, j);
                v[j] = x;
                float s = 0.0f, c = 1.0f;
#pragma unroll
                for (int k = 0; k < NFX; ++k) {
                    pe_band(x, k, &s, &c);
                    v[XD + k * 2 * XD + j] = s;
                    v[XD + k * 2 * XD + XD + j] = c;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < KPE / 8; ++c)
            out[c * kTileM] = make_uint4(pack_h2(v[8 * c], v[8 * c + 1]), pack_h2(v[8 * c + 2], v[8 * c + 3]),
                                         pack_h2(v[8 * c + 4], v[8 * c + 5]), pack_h2(v[8 * c + 6], v[8 * c + 7]));
    }
    if (KAUX > 0) {
        float v[KAUX > 0 ? KAUX : 1];
#pragma unroll
        for (int c = 0; c < KAUX; ++c) v[c] = 0.0f;
        if (row >= 0 && !a.sigma_only) {
            if (NFD > 0) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float d = a.src.dir(row, j);
                    v[j] = d;
                    float s = 0.0f, c = 1.0f;
#pragma unroll
                    for (int k = 0; k < NFD; ++k) {
                        pe_band(d, k, &s, &c);
                        v[3 + k * 6 + j] = s;
                        v[3 + k * 6 + 3 + j] = c;
                    }
                }
            }
            if (APP > 0) {
                int sub = a.fixed_sub;
                if (a.counters) {
                    sub = 0;
                    while (sub + 1 < a.n_sub && slot0 >= a.counters[CNT_START + sub + 1]) ++sub;
                }
                int id = (int)a.src.index(row);
                id = min(max(id, 0), a.nd.app_count - 1);
                const float4* e4 = reinterpret_cast<const float4*>(a.packed + (size_t)sub * a.lay.total + a.lay.emb + (size_t)id * APP);
#pragma unroll
                for (int j = 0; j < APP / 4; ++j) {
                    const float4 e = __ldg(e4 + j);
                    v[IN_DIR + 4 * j + 0] = e.x;
                    v[IN_DIR + 4 * j + 1] = e.y;
                    v[IN_DIR + 4 * j + 2] = e.z;
                    v[IN_DIR + 4 * j + 3] = e.w;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < KAUX / 8; ++c)
            out[(KPE / 8 + c) * kTileM] = make_uint4(pack_h2(v[8 * c], v[8 * c + 1]), pack_h2(v[8 * c + 2], v[8 * c + 3]),
                                                     pack_h2(v[8 * c + 4], v[8 * c + 5]), pack_h2(v[8 * c + 6], v[8 * c + 7]));
    }
}

// ---- lean primitives for the MMA-issuing warp (raw shared-memory addresses, no per-slab descriptor rebuild) ----
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("mn_mlp_tc: mbarrier timeout (mma) block %d\n", (int)blockIdx.x);
            __trap();
        }
    }
}
// Non-blocking look-ahead probe (mbarrier.test_wait): issued BEFORE the current stage's work so that its ~100-cycle latency
// overlaps that work instead of heading the next stage's dependent chain.  A single thread pays ~260 cycles per ring stage
// for wait -> issue -> signal when these are strictly sequential (scripts/probes/l2_stream_probe.cu: throughput of a
// one-thread TMA ring is proportional to the stage size and independent of its depth) - as long as two 128-cycle MMAs.
__device__ __forceinline__ uint32_t mbar_test_a(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    return ok;
}
// one ring stage worth of MMAs (one or two K=16 steps) + the commit that releases the stage, issued by one
// elected lane; everything is predicated, no branches.
__device__ __forceinline__ void mma_stage(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint64_t ad2, uint64_t bd2, uint32_t idesc,
                                          uint32_t accum, uint32_t two, uint32_t empty_bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e, p, q;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.and.b32 q, %7, 0, e;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(bd), "l"(ad2), "l"(bd2), "r"(idesc), "r"(accum), "r"(two), "r"(empty_bar_addr)
        : "memory");
}
// Two consecutive ring stages of one GEMM in one issue block: up to four K=16 MMAs (A contiguous in the activation buffer,
// B in stage a then stage b), each stage released by its own commit.  n_b = MMAs of the second stage (1 or 2).
__device__ __forceinline__ void mma_stage2(uint32_t d_tmem, uint64_t ad, uint64_t a_step, uint64_t bda, uint64_t bdb, uint64_t b_step,
                                           uint32_t idesc, uint32_t accum, uint32_t n_b, uint32_t empty_a, uint32_t empty_b) {
    asm volatile(
        "{\n\t.reg .pred e, p, q;\n\t.reg .b64 a1, a2, a3, b1, b3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.gt.and.u32 q, %8, 1, e;\n\t"
        "add.u64 a1, %1, %2;\n\tadd.u64 a2, a1, %2;\n\tadd.u64 a3, a2, %2;\n\t"
        "add.u64 b1, %3, %5;\n\tadd.u64 b3, %4, %5;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %6, 1;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, %4, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %6, 1;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(a_step), "l"(bda), "l"(bdb), "l"(b_step), "r"(idesc), "r"(accum), "r"(n_b), "r"(empty_a), "r"(empty_b)
        : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar_addr) : "memory");
}


// up to four K=16 steps against one ring stage (A from shared memory: descriptor + a_step per step); one elected lane issues,
// then releases the stage.
__device__ __forceinline__ void ts_stage_smem(uint32_t d_tmem, uint64_t ad, uint64_t a_step, uint64_t bd, uint64_t b_step,
                                              uint32_t idesc, uint32_t accum, int nk, uint32_t empty_bar) {
    asm volatile(
        "{\n\t.reg .pred e, p, q1, q2, q3;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.gt.and.s32 q1, %7, 1, e;\n\t"
        "setp.gt.and.s32 q2, %7, 2, e;\n\t"
        "setp.gt.and.s32 q3, %7, 3, e;\n\t"
        "add.u64 a1, %1, %2;\n\tadd.u64 a2, a1, %2;\n\tadd.u64 a3, a2, %2;\n\t"
        "add.u64 b1, %3, %4;\n\tadd.u64 b2, b1, %4;\n\tadd.u64 b3, b2, %4;\n\t"
        "@e  tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, p;\n\t"
        "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %5, 1;\n\t"
        "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %5, 1;\n\t"
        "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %5, 1;\n\t"
        "@e  tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
        ::"r"(d_tmem), "l"(ad), "l"(a_step), "l"(bd), "l"(b_step), "r"(idesc), "r"(accum), "r"(nk), "r"(empty_bar)
        : "memory");
}


// ------------------------------------------------------------------------------------------------
// the MLP kernel
// ------------------------------------------------------------------------------------------------
struct TcArgs {
    MlpArgs m;
    TcPlan plan;
    const unsigned char* wpack;   // per sub-module: [hi plane][lo plane?][fp32 block]
    const __half* ximg;           // feature tiles (hi plane; lo plane at +x_plane_halves)
    int64_t x_plane_halves;
    int split;                    // 1: three MMA passes (hi*hi + hi*lo + lo*hi)
    int desc_swap;                // debug: 1 = record the in-kernel timeline (MN_TC_TRACE)
    int64_t n_tiles_cap;
    // ---- training (tc_f16 training path, mn_train_tc.cuh).  Tapes hold, per 128-slot tile, fp16 tile images in the layout of
    // the activation buffer ([cols/8][128][8]): activations H_0 .. H_{layers-1}, F (xyz_encoding_final), G (dir_a_encoding).
    unsigned char* tape_act;      // PP_TRAIN_FWD: written;  PP_DGRAD: read (ReLU masks)
    float* tape_f32;              // per tile [4][128]: sigma pre-activation, rgb (3)     (written / read)
    unsigned char* tape_dz;       // PP_DGRAD: gradient images dZ_0 .. dZ_{layers-1}, dZ_final, dZ_dira (same layout, scaled fp16)
    float* tape_gf32;             // PP_DGRAD: per tile [4][128]: d sigma pre-activation, d rgb pre-activation (3), UNscaled fp32
    const float* grad_out;        // PP_DGRAD: [rows][rgb_dim + 1] upstream gradient
    float* emb_sum;               // PP_DGRAD: [n_sub][app_count][L/2] per-image sums of dZ_dira rows (appearance-embedding gradient)
    const float* scale;           // PP_DGRAD: device scalar S (power of two): gradient images hold S * dZ
    int64_t act_tile_bytes;       // bytes of one tile's record in tape_act / tape_dz
    int layers;
    // ---- TMEM ping-pong kernel (mn_mlp_tp.cuh): half-major weight images and the two role tables
    const unsigned char* tpack;
    size_t tp_sub_bytes;
    const uint4* tp_prog;
    int tp_n[4];
};

struct SmemLayout {
    // byte offsets inside dynamic shared memory
    int ring, h, xa, f32, sigp, bars, total, stages;
};

constexpr int kSmemMax = 227 * 1024;

__host__ __device__ inline SmemLayout smem_layout(const TcPlan& p, bool split) {
    SmemLayout s;
    const int kx = p.kpe > p.kaux ? p.kpe : p.kaux;
    const int fixed = p.L * kTileM * 2 * (split ? 2 : 1) + kx * kTileM * 2 + ((p.f32_floats * 4 + 15) / 16) * 16 + 2048 + 256;
    int st = (kSmemMax - fixed) / ring_stage_bytes(split);
    if (st > kMaxStages) st = kMaxStages;
    s.stages = st;
    s.ring = 0;
    s.h = s.ring + st * ring_stage_bytes(split);
    s.xa = s.h + p.L * kTileM * 2 * (split ? 2 : 1);   // split: hi plane then lo plane
    s.f32 = s.xa + kx * kTileM * 2;
    s.sigp = s.f32 + ((p.f32_floats * 4 + 15) / 16) * 16;
    s.bars = s.sigp + 2048;
    s.total = s.bars + 256;
    return s;
}

// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue.  Epilogue warp w owns TMEM lanes
// 32*(w%4).. (hardware rule) and, within every 64-column slab of the accumulator, the 32-column half
// (w-2)/4; a slab of the next layer's A operand is published (mbarrier hready[slab]) as soon as all eight
// warps have written their part, so the next layer's MMAs trail the epilogue slab by slab while the
// other accumulator buffer is still being drained.
template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 1) tc_mlp_kernel(const TcArgs A) {
    constexpr int kSlabCols = ring_slab_cols(kSplit);
    constexpr int kStageBytes = ring_stage_bytes(kSplit);
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const SmemLayout SL = smem_layout(P, kSplit);
    const int kStages = SL.stages;
    unsigned char* ring = smem + SL.ring;
    unsigned char* Hs = smem + SL.h;
    unsigned char* XA = smem + SL.xa;
    float* F32 = reinterpret_cast<float*>(smem + SL.f32);
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);   // [4][128] partial sigma dot products
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;                  // [kMaxStages]
    uint64_t* empty = bars + kMaxStages;    // [kMaxStages]
    uint64_t* xa_full = bars + 2 * kMaxStages;
    uint64_t* xa_empty = xa_full + 1;
    uint64_t* acc_full = xa_full + 2;       // [2]
    uint64_t* hready = xa_full + 4;         // [4]
    uint64_t* f32_full = xa_full + 8;
    uint64_t* f32_empty = xa_full + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xa_full + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(xa_full, 1);
        mbar_init(xa_empty, 1);
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&hready[i], kEpiWarps);
        mbar_init(f32_full, 1);
        mbar_init(f32_empty, kEpiWarps);
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

    auto sub_of = [&](int64_t tile) -> int {
        int sub = A.m.fixed_sub;
        if (A.m.counters) {
            sub = 0;
            const int64_t s0 = tile * kTileM;
            while (sub + 1 < A.m.n_sub && s0 >= A.m.counters[CNT_START + sub + 1]) ++sub;
        }
        return sub;
    };
    constexpr int npass = kSplit ? 3 : 1;

    if (warp == kWarpProd) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, xphase = 0, fphase = 0;
            const uint32_t f32_bytes = (uint32_t)(((P.f32_floats * 4 + 15) / 16) * 16);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int sub = sub_of(tile);
                const unsigned char* wsub = A.wpack + (size_t)sub * P.sub_bytes;
                const size_t f32_off = (size_t)P.plane_bytes * 2;   // both planes are always packed
                // biases + sigma weights of this tile's sub-module
                mbar_wait(f32_empty, fphase ^ 1);
                mbar_expect_tx(f32_full, f32_bytes);
                bulk_g2s(F32, wsub + f32_off, f32_bytes, f32_full);
                fphase ^= 1;
                for (int gi = 0; gi < n_gemm; ++gi) {
                    const TcGemm& g = P.g[gi];
                    for (int pass = 0; pass < npass; ++pass) {
                        // pass 0: A_hi * W_hi, pass 1: A_hi * W_lo, pass 2: A_lo * W_hi
                        const unsigned char* wimg = wsub + (pass == 1 ? (size_t)P.plane_bytes : 0) + g.w_off;
                        int kbase = 0;
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            if (g.src[sgi] != SRC_H) {
                                const __half* xt = A.ximg + (pass == 2 ? A.x_plane_halves : 0) +
                                                   tile * (int64_t)(P.kpe + P.kaux) * kTileM +
                                                   (g.src[sgi] == SRC_XAUX ? (int64_t)P.kpe * kTileM : 0);
                                mbar_wait(xa_empty, xphase ^ 1);
                                mbar_expect_tx(xa_full, (uint32_t)(kseg * kTileM * 2));
                                bulk_g2s(XA, xt, (uint32_t)(kseg * kTileM * 2), xa_full);
                                xphase ^= 1;
                            }
                            for (int k0 = 0; k0 < kseg; k0 += kSlabCols) {
                                const int kc = min(kSlabCols, kseg - k0);
                                const uint32_t bytes = (uint32_t)(kc * g.n * 2);
                                mbar_wait(&empty[stage], phase ^ 1);
                                mbar_expect_tx(&full[stage], bytes);
                                bulk_g2s(ring + (size_t)stage * kStageBytes, wimg + (size_t)(kbase + k0) * g.n * 2, bytes,
                                         &full[stage]);
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                            }
                            kbase += kseg;
                        }
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer ===========================
        // The whole warp runs the loop (so addresses/descriptors stay in uniform registers); one elected lane
        // issues tcgen05.mma / tcgen05.commit.
        {
            int stage = 0;
            uint32_t phase = 0, xphase = 0, gidx = 0;
            uint32_t hph0 = 0, hph1 = 0, hph2 = 0, hph3 = 0;
            const uint32_t h_base = smem_u32(Hs), xa_base = smem_u32(XA), ring_base = smem_u32(ring);
            const uint64_t a_step = (uint64_t)((2 * kTileM * 16) >> 4);           // +16 K-columns of A (descriptor address field)
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int gi = 0; gi < n_gemm; ++gi, ++gidx) {
                    const TcGemm& g = P.g[gi];
                    const uint32_t idesc = make_idesc(g.n);
                    const uint32_t d_tmem = tmem_base + (gidx & 1u) * 256u;
                    const uint64_t b_step = (uint64_t)((2 * g.n * 16) >> 4);      // +16 K-columns of B
                    uint32_t accum = 0;
                    for (int pass = 0; pass < npass; ++pass) {
                        for (int sgi = 0; sgi < g.nseg; ++sgi) {
                            const int kseg = g.k[sgi];
                            const bool from_x = g.src[sgi] != SRC_H;
                            // lo plane of H lives right after the hi plane in the H buffer (split mode)
                            const uint32_t a_base = from_x ? xa_base : h_base + (pass == 2 ? (uint32_t)(P.L * kTileM * 2) : 0u);
                            if (from_x) {
                                mbar_wait(xa_full, xphase);
                                xphase ^= 1;
                            }
                            for (int k0 = 0; k0 < kseg; k0 += kSlabCols) {
                                const int kc = min(kSlabCols, kseg - k0);
                                if (!from_x && pass == 0 && (k0 & 63) == 0) {
                                    // the previous GEMM's epilogue has published this 64-column slab of H
                                    const int hs = k0 >> 6;
                                    if (hs == 0) { mbar_wait(&hready[0], hph0); hph0 ^= 1; }
                                    else if (hs == 1) { mbar_wait(&hready[1], hph1); hph1 ^= 1; }
                                    else if (hs == 2) { mbar_wait(&hready[2], hph2); hph2 ^= 1; }
                                    else { mbar_wait(&hready[3], hph3); hph3 ^= 1; }
                                }
                                mbar_wait(&full[stage], phase);
                                tc_fence_after();
                                uint64_t ad = make_desc(a_base + (uint32_t)(k0 / 8) * (kTileM * 16), kTileM * 16, 128);
                                uint64_t bd = make_desc(ring_base + (uint32_t)stage * kStageBytes, (uint32_t)g.n * 16, 128);
                                // up to four MMAs and the commit that frees the ring stage, one predicated straight-line block
                                ts_stage_smem(d_tmem, ad, a_step, bd, b_step, idesc, accum, kc >> 4, smem_u32(&empty[stage]));
                                accum = 1;
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                            }
                            if (from_x && elect_one()) tc_commit(xa_empty);
                            __syncwarp();
                        }
                    }
                    if (elect_one()) tc_commit(&acc_full[gidx & 1u]);
                    __syncwarp();
                }
            }
        }
    } else {
        // =========================== epilogue (16 warps) ===========================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int part = warp >> 2;            // which 16-column piece of every 64-column slab
        const int r = q * 32 + lane;                 // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t acc_phase0 = 0, acc_phase1 = 0, fphase = 0, gidx = 0;
        const int L = P.L;
        const size_t lo_off = (size_t)L * kTileM * 2;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t slot = tile * kTileM + r;
            int64_t row = -1;
            if (slot < n_slots) row = A.m.slot_row ? (int64_t)A.m.slot_row[slot] : slot;
            mbar_wait(f32_full, fphase);
            fphase ^= 1;
            float sigma = 0.0f;
            for (int gi = 0; gi < n_gemm; ++gi, ++gidx) {
                const TcGemm& g = P.g[gi];
                const uint32_t ab = gidx & 1u;
                if (ab) { mbar_wait(&acc_full[1], acc_phase1); acc_phase1 ^= 1; }
                else    { mbar_wait(&acc_full[0], acc_phase0); acc_phase0 ^= 1; }
                tc_fence_after();
                const uint32_t t_acc = t_lane + ab * 256u;
                const float* bias = F32 + g.bias_off;
                if (g.epi == EPI_RGB) {
                    if (part == 0) {
                        uint32_t v[32];
                        tmem_ld32(t_acc, v);
                        tmem_ld_wait();
                        if (row >= 0) tc_emit_rgb(A.m, A.m.nd.affine ? sub_of(tile) : 0, row, slot, v, bias, sigma);
                    }
                    tc_fence_before();
                } else {
                    const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                    const bool publish = !(want_sigma && A.m.sigma_only);   // nobody reads H after the last trunk layer
                    const float* sw = F32 + P.sigma_w_off;
                    float sacc = 0.0f;
                    const int nslab = (g.n + 63) >> 6;
                    for (int j = 0; j < nslab; ++j) {
                        const int c0 = 64 * j + 16 * part;
                        if (c0 < g.n) {
                            unsigned char* dst = Hs + (size_t)(c0 >> 3) * (kTileM * 16) + (size_t)r * 16;
                            if (g.epi == EPI_RELU)
                                epi_piece16<kSplit, true, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, lo_off, true);
                            else if (g.epi == EPI_LINEAR)
                                epi_piece16<kSplit, false, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, lo_off, true);
                            else
                                sacc += epi_piece16<kSplit, true, true>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, lo_off, publish);
                            if (publish) fence_proxy_async();   // generic-proxy stores to H -> visible to the tensor core
                        }
                        if (publish) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&hready[j]);
                        }
                    }
                    if (want_sigma) {
                        SIGP[part * kTileM + r] = sacc;
                        asm volatile("bar.sync 1, 512;" ::: "memory");
                        if (part == 0) {
                            // sigma bias is stored right after sigma_w
                            float s = ((SIGP[r] + SIGP[kTileM + r]) + (SIGP[2 * kTileM + r] + SIGP[3 * kTileM + r])) + sw[L];
                            if (A.m.sigma_noise && row >= 0) s = s + A.m.sigma_noise[row];
                            sigma = A.m.nd.softplus ? mn_softplus_shifted(s) : fmaxf(s, 0.0f);
                            if (A.m.sigma_only && row >= 0) {
                                const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                                A.m.out[o] = A.m.slot_w ? sigma * A.m.slot_w[slot] : sigma;
                            }
                        }
                        tc_fence_before();
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(f32_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpProd) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}



// ------------------------------------------------------------------------------------------------
// Ping-pong variant (single-pass fp16): two 128-row tiles per CTA, each with its own activation buffer
// and TMEM accumulator.  GEMMs are issued X_l, Y_l, X_l+1, Y_l+1, ...: the epilogue of X_l (16 warps)
// runs entirely under the MMAs of Y_l and vice versa, so the tensor pipe only idles at pipeline fill.
// ------------------------------------------------------------------------------------------------
// A ring stage (16 KiB) carries either 32 K-columns of weights (activation segment: A operand = the tile's H buffer,
// two K=16 MMAs) or, for the feature segments (PE / direction+appearance tiles), 16 K-columns of weights (<= 8 KiB) PLUS
// the matching 16 K-columns x 128 rows of the tile's feature image (4 KiB at +8 KiB, one MMA): the features travel in
// the same stream as the weights, so there is no separate feature buffer and no producer stall waiting for it.
constexpr int kPPMaxStages = 8;
constexpr int kPPSlabCols = 32;
constexpr int kPPStageBytes = kPPSlabCols * 256 * 2;
constexpr int kPPXCols = 16;
constexpr int kPPXOff = 8192;

struct PPLayout {
    int ring, h, f32, f32_stride, sigp, bars, prog, total, stages;
};
// Stage program of the ping-pong kernel: ONE 16-byte entry per ring stage of a tile pair, shared by the TMA producer and the
// MMA issuer, so that both single-thread roles run a flat loop (no nested GEMM / segment / K loops, no address arithmetic
// beyond one add): their per-stage dependent chain is what paces the kernel (see the probe notes at mbar_test_a).
//   x = weight byte offset inside the sub-module image   y = weight bytes | (feature offset / 16, 0xFFFF = none) << 16
//   z = flags | MMA N << 8                               w = activation-operand offset >> 4 (descriptor units)
constexpr int kPPMaxProg = 208;
enum { PF_SLOT1 = 1, PF_FIRST = 2, PF_LAST = 4, PF_FROM_X = 8, PF_TWO = 16, PF_PAIR = 32 };   // PF_PAIR: this and the next entry are consecutive activation stages of one GEMM - the MMA warp issues them together

__host__ __device__ inline PPLayout pp_layout(const TcPlan& p) {
    PPLayout s;
    s.f32_stride = ((p.f32_floats * 4 + 15) / 16) * 16;
    const int fixed = 2 * p.L * kTileM * 2 + s.f32_stride + 2048 + 256 + kPPMaxProg * 16;
    int kPPStages = (kSmemMax - fixed) / kPPStageBytes;
    if (kPPStages > kPPMaxStages) kPPStages = kPPMaxStages;
    s.stages = kPPStages;
    s.ring = 0;
    s.h = kPPStages * kPPStageBytes;
    s.f32 = s.h + 2 * p.L * kTileM * 2;
    s.sigp = s.f32 + s.f32_stride;    // ONE fp32 block: both tiles of a pair belong to the same sub-module
    s.bars = s.sigp + 2048;
    s.prog = s.bars + 256;
    s.total = s.prog + kPPMaxProg * 16;
    return s;
}

__host__ __device__ inline int pp_gemm_stages(const TcGemm& g) {
    int n = 0;
    for (int sgi = 0; sgi < g.nseg; ++sgi)
        n += g.src[sgi] != SRC_H ? (g.k[sgi] + kPPXCols - 1) / kPPXCols : (g.k[sgi] + kPPSlabCols - 1) / kPPSlabCols;
    return n;
}
__host__ __device__ inline int pp_prog_entries(const TcPlan& p, int n_gemm) {
    int n = 0;
    for (int gi = 0; gi < n_gemm; ++gi) n += 2 * pp_gemm_stages(p.g[gi]);
    return n;
}

// Ping-pong kernel (see the header of this section).  Warps 0..15 epilogue, 16 TMA producer, 17 MMA issuer.
template <int kMode>
__global__ void __launch_bounds__(kPPThreads, 1) tc_mlp_pp_kernel(const TcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcPlan& P = A.plan;
    const PPLayout SL = pp_layout(P);
    const int kPPStages = SL.stages;
    unsigned char* ring = smem + SL.ring;
    unsigned char* Hs = smem + SL.h;
    float* F32 = reinterpret_cast<float*>(smem + SL.f32);
    float* SIGP = reinterpret_cast<float*>(smem + SL.sigp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* full = bars;            // [<=8]
    uint64_t* empty = bars + 8;       // [<=8]
    uint64_t* acc_full = bars + 16;   // [2] per tile slot
    uint64_t* epi_done = bars + 18;   // [2]
    uint64_t* f32_full = bars + 20;   // [2]
    uint64_t* f32_empty = bars + 22;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_slots = A.m.counters ? A.m.counters[CNT_NSLOTS] : A.m.B;
    const int64_t n_tiles = (n_slots + kTileM - 1) / kTileM;
    const int n_gemm = A.m.sigma_only ? P.n_trunk : P.n_gemm;
    const int h_bytes = P.L * kTileM * 2;
    const float* SW = F32 + P.sigma_w_off;
    const float* RGBB = F32 + P.g[P.n_gemm - 1].bias_off;

    // ---- stage program (see kPPMaxProg): entries in the order both roles walk a tile pair: GEMM, slot, segment, K
    uint4* PROG = reinterpret_cast<uint4*>(smem + SL.prog);
    auto stages_of = [&](const TcGemm& g) -> int { return pp_gemm_stages(g); };
    const int n_prog = pp_prog_entries(P, n_gemm);       // <= kPPMaxProg: checked by the launcher
    if (threadIdx.x >= 64 && threadIdx.x < 66) PROG[n_prog + (threadIdx.x - 64)] = make_uint4(0u, 0u, 0u, 0u);   // read-ahead padding
    if ((int)threadIdx.x < 2 * n_gemm) {
        const int gi = threadIdx.x >> 1, sl = threadIdx.x & 1;
        const TcGemm& g = P.g[gi];
        int e = 0;
        for (int j = 0; j < gi; ++j) e += 2 * stages_of(P.g[j]);
        const int total = stages_of(g);
        e += sl * total;
        int kbase = 0, cnt = 0;
        for (int sgi = 0; sgi < g.nseg; ++sgi) {
            const bool fx = g.src[sgi] != SRC_H;
            const int kk = g.k[sgi];
            const int step = fx ? kPPXCols : kPPSlabCols;
            for (int k0 = 0; k0 < kk; k0 += step, ++cnt, ++e) {
                const int kc = min(step, kk - k0);
                const uint32_t xo = fx ? (uint32_t)(((g.src[sgi] == SRC_XAUX ? P.kpe * kTileM * 2 : 0) + k0 * kTileM * 2) >> 4) : 0xFFFFu;
                const bool pair = !fx && ((k0 / kPPSlabCols) & 1) == 0 && k0 + kPPSlabCols < kk;      // even activation stage with a successor
                const uint32_t fl = (sl ? PF_SLOT1 : 0) | (cnt == 0 ? PF_FIRST : 0) | (cnt == total - 1 ? PF_LAST : 0) | (fx ? PF_FROM_X : 0) |
                                    ((!fx && kc == kPPSlabCols) ? PF_TWO : 0) | (pair ? PF_PAIR : 0);
                PROG[e] = make_uint4((uint32_t)(g.w_off + (kbase + k0) * g.n * 2), (uint32_t)(kc * g.n * 2) | (xo << 16),
                                     fl | ((uint32_t)g.n << 8), fx ? 0u : (uint32_t)(((k0 >> 3) * (kTileM * 16)) >> 4));
            }
            kbase += kk;
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPPMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&epi_done[i], kEpiWarps);
            mbar_init(&f32_full[i], 1);
            mbar_init(&f32_empty[i], kEpiWarps);
        }
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

    if (warp == kWarpProd) {
        // =========================== TMA producer: table-driven (PROG) ===========================
        // One thread.  Its per-stage dependent chain (wait -> expect_tx -> copy) is as long as the stage's two MMAs, so the
        // loop is kept flat: one 16-byte table entry per ring stage, look-ahead probe of the next stage's empty barrier.
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, fph_e = 0, ahead = 0;
            int last_sub = -1;
            const uint32_t f32_bytes = (uint32_t)SL.f32_stride;
            const uint32_t empty_pa = smem_u32(empty), full_pa = smem_u32(full), ring_pa = smem_u32(ring);
            const uint32_t nst = (uint32_t)kPPStages;
            const int64_t xtile_bytes = (int64_t)(P.kpe + P.kaux) * kTileM * 2;
            for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
                const int64_t t0 = 2 * pr;
                const int sub0 = sub_of(t0);
                const unsigned char* wsub = A.wpack + (size_t)sub0 * P.sub_bytes;
                const unsigned char* xt0 = reinterpret_cast<const unsigned char*>(A.ximg) + t0 * xtile_bytes;
                const bool valid1 = t0 + 1 < n_tiles;
                if (sub0 != last_sub) {
                    // the bias / sigma block is re-staged only when the sub-module changes (a handful of times per launch):
                    // the weight stream of consecutive pairs is not interrupted by waiting for the epilogue
                    if (last_sub >= 0) { mbar_wait(&f32_empty[0], fph_e); fph_e ^= 1; }
                    const unsigned char* fsrc = wsub + (size_t)P.f32_off;
                    mbar_expect_tx(&f32_full[0], f32_bytes);
                    bulk_g2s(reinterpret_cast<unsigned char*>(F32), fsrc, f32_bytes, &f32_full[0]);
                    last_sub = sub0;
                }
                for (int e = 0; e < n_prog; ++e) {
                    const uint4 E = PROG[e];
                    const bool slot1 = (E.z & PF_SLOT1) != 0;
                    if (slot1 && !valid1) continue;
                    const uint32_t cur = stage;
                    if (!ahead) mbar_wait_a(empty_pa + 8u * cur, phase ^ 1);
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                    ahead = mbar_test_a(empty_pa + 8u * stage, phase ^ 1);       // next stage: latency overlaps the copies below
                    const uint32_t bar = full_pa + 8u * cur, dst = ring_pa + cur * (uint32_t)kPPStageBytes;
                    const uint32_t wbytes = E.y & 0xFFFFu, xo = E.y >> 16;
                    const bool has_x = xo != 0xFFFFu;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(wbytes + (has_x ? (uint32_t)(kPPXCols * kTileM * 2) : 0u)) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(wsub + E.x), "r"(wbytes), "r"(bar) : "memory");
                    if (has_x)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst + (uint32_t)kPPXOff), "l"(xt0 + (slot1 ? xtile_bytes : 0) + (int64_t)(xo << 4)),
                                       "r"((uint32_t)(kPPXCols * kTileM * 2)), "r"(bar) : "memory");
                }
            }
        }
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer (whole warp, one elected lane issues): table-driven (PROG) ===========================
        uint32_t stage = 0, phase = 0, eph0 = 0, eph1 = 0;
        uint32_t ahead = 0;              // the look-ahead probe of the CURRENT stage's full barrier already succeeded
        // data-gradient mode: the first GEMM of every pair waits for the head stage the epilogue warps run first
        bool started0 = kMode == PP_DGRAD, started1 = kMode == PP_DGRAD;
        const uint32_t h_base = smem_u32(Hs), ring_base = smem_u32(ring);
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const uint32_t acc_full_a = smem_u32(acc_full), epi_done_a = smem_u32(epi_done);
        const uint32_t nst = (uint32_t)kPPStages;
        const uint64_t xd0 = make_desc(ring_base + (uint32_t)kPPXOff, kTileM * 16, 128);
        const uint64_t bd_base = make_desc(ring_base, 0, 128);                 // LBO (= N * 16 bytes) is added per entry
        const uint64_t hd0 = make_desc(h_base, kTileM * 16, 128), hd1 = make_desc(h_base + (uint32_t)h_bytes, kTileM * 16, 128);
        const uint64_t a_step = (uint64_t)((2 * kTileM * 16) >> 4);
        const uint64_t st_step = (uint64_t)(kPPStageBytes >> 4);
        uint32_t accum = 0;
        for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
            const bool valid1 = 2 * pr + 1 < n_tiles;
            uint4 E = PROG[0];
            for (int e = 0; e < n_prog;) {
                const uint32_t fl = E.z & 0xFFu, N = E.z >> 8;
                const uint4 E1 = PROG[e + 1], E2 = PROG[e + 2];      // prefetched: the table has two spare entries at its end
                if ((fl & PF_SLOT1) && !valid1) { ++e; E = E1; continue; }
                const uint32_t sl = (fl & PF_SLOT1) ? 1u : 0u;
                if (fl & PF_FIRST) {
                    // the previous GEMM of this tile slot has been drained from TMEM and its activations are in H[sl]
                    if (sl == 0) { if (started0) { mbar_wait_a(epi_done_a, eph0); eph0 ^= 1; } started0 = true; }
                    else         { if (started1) { mbar_wait_a(epi_done_a + 8, eph1); eph1 ^= 1; } started1 = true; }
                    accum = 0;
                }
                const uint32_t cur = stage;
                if (!ahead) mbar_wait_a(full_a + 8u * cur, phase);
                if (++stage == nst) { stage = 0; phase ^= 1; }
                if (fl & PF_PAIR) {
                    // two activation stages (up to 4 MMAs) per iteration: the loop overhead of this single warp - not the data -
                    // is what paces the kernel (ncu: it never waits on `full`, it is busy with ~90 instructions per stage)
                    const uint32_t cur1 = stage;
                    if (!mbar_test_a(full_a + 8u * cur1, phase)) mbar_wait_a(full_a + 8u * cur1, phase);
                    tc_fence_after();
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                    ahead = mbar_test_a(full_a + 8u * stage, phase);
                    const uint64_t ad = (sl ? hd1 : hd0) + (uint64_t)E.w;
                    const uint64_t bb = bd_base + ((uint64_t)N << 16);
                    mma_stage2(tmem_base + sl * 256u, ad, a_step, bb + (uint64_t)cur * st_step, bb + (uint64_t)cur1 * st_step, (uint64_t)(2u * N),
                               make_idesc((int)N), accum, ((E1.z & PF_TWO) ? 2u : 1u), empty_a + 8u * cur, empty_a + 8u * cur1);
                    accum = 1;
                    if (E1.z & PF_LAST) commit_elect(acc_full_a + 8u * sl);
                    e += 2;
                    E = E2;
                    continue;
                }
                tc_fence_after();
                ahead = mbar_test_a(full_a + 8u * stage, phase);              // next stage, overlapped with the issue below
                const uint64_t so = (uint64_t)cur * st_step;
                const uint64_t ad = (fl & PF_FROM_X) ? xd0 + so : (sl ? hd1 : hd0) + (uint64_t)E.w;
                const uint64_t bd = bd_base + so + ((uint64_t)N << 16);        // LBO field = N * 16 bytes >> 4 = N
                mma_stage(tmem_base + sl * 256u, ad, bd, ad + a_step, bd + (uint64_t)(2u * N), make_idesc((int)N), accum,
                          (fl & PF_TWO) ? 1u : 0u, empty_a + 8u * cur);
                accum = 1;
                if (fl & PF_LAST) commit_elect(acc_full_a + 8u * sl);
                ++e;
                E = E1;
            }
        }
    } else {
        // =========================== epilogue (16 warps) ===========================
        const int q = warp & 3;
        const int part = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t aph0 = 0, aph1 = 0, fph0 = 0;
        int last_sub = -1;
        const int L = P.L;
        for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
            const int64_t t0 = 2 * pr;
            const bool valid1 = t0 + 1 < n_tiles;
            {
                const int sub0 = sub_of(t0);
                if (sub0 != last_sub) {
                    if (last_sub >= 0) {            // done with the previous sub-module's block
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&f32_empty[0]);
                    }
                    mbar_wait(&f32_full[0], fph0);
                    fph0 ^= 1;
                    last_sub = sub0;
                }
            }
            int64_t slot_[2], row_[2] = {-1, -1};
            float sigma_[2] = {0.0f, 0.0f};
            for (int sl = 0; sl < 2; ++sl) {
                if (sl == 1 && !valid1) continue;
                slot_[sl] = (t0 + sl) * kTileM + r;
                if (slot_[sl] < n_slots) row_[sl] = A.m.slot_row ? (int64_t)A.m.slot_row[slot_[sl]] : slot_[sl];
            }
            float dsig_[2] = {0.0f, 0.0f};       // PP_DGRAD: S * d(sigma pre-activation) of this thread's row, per tile slot
            if (kMode == PP_DGRAD) {
                // ---- head stage of the data-gradient chain (nerf.py:132-160 backwards): upstream gradient x blend weight ->
                // sigmoid' / softplus' -> rgb Linear transposed (3 -> L/2, CUDA cores) -> ReLU mask of dir_a_encoding -> dZ_dira
                // as the first A operand (columns 0 .. L/2-1 of the activation buffer) and on the gradient tape; per-image sums
                // of its rows for the appearance-embedding gradient; head pre-activation gradients in fp32 for their own Linears.
                const float S = *A.scale;
                const int sub0 = last_sub;
                const float* Wr = F32 + L;                                  // [3][L/2] rgb weights (fp32 block of the data-gradient plan)
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl == 1 && !valid1) continue;
                    const int64_t tile = t0 + sl;
                    const int64_t row = row_[sl];
                    const float* tf = A.tape_f32 + (size_t)tile * 5 * kTileM + r;
                    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
                    if (row >= 0) {
                        const float4 g = *reinterpret_cast<const float4*>(A.grad_out + row * 4);
                        const float w = A.m.slot_w ? A.m.slot_w[slot_[sl]] : 1.0f;
                        g0 = g.x * w; g1 = g.y * w; g2 = g.z * w; g3 = g.w * w;
                    }
                    const float c0v = tf[1 * kTileM], c1v = tf[2 * kTileM], c2v = tf[3 * kTileM], pre = tf[0];
                    const float d0 = (g0 * (1.0f - c0v)) * c0v, d1 = (g1 * (1.0f - c1v)) * c1v, d2 = (g2 * (1.0f - c2v)) * c2v;
                    float dsp;
                    if (A.m.nd.softplus) { const float y = pre - 1.0f; dsp = y > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-y)); }
                    else dsp = pre > 0.0f ? 1.0f : 0.0f;
                    const float ds = g3 * dsp;
                    dsig_[sl] = ds * S;
                    if (part == 0) {
                        float* tg = A.tape_gf32 + (size_t)tile * 4 * kTileM + r;
                        tg[0] = ds; tg[1 * kTileM] = d0; tg[2 * kTileM] = d1; tg[3 * kTileM] = d2;
                    }
                    const int id = (int)tf[4 * kTileM];
                    const unsigned char* gimg = A.tape_act + (size_t)tile * A.act_tile_bytes + (size_t)(A.layers + 1) * L * kTileM * 2;
                    unsigned char* dimg = A.tape_dz + (size_t)tile * A.act_tile_bytes + (size_t)(A.layers + 1) * L * kTileM * 2;
                    unsigned char* Hsl = Hs + (size_t)sl * h_bytes;
                    const int half = L / 2, per = half / 4;                 // columns of dZ_dira handled by this thread (32 for L = 256)
                    for (int kk = 0; kk < per; kk += 8) {
                        const int k0 = part * per + kk;
                        const uint4 gm = *reinterpret_cast<const uint4*>(gimg + (size_t)(k0 >> 3) * (kTileM * 16) + (size_t)r * 16);
                        const __half2* gh = reinterpret_cast<const __half2*>(&gm);
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int k = k0 + e;
                            float acc = Wr[k] * d0;
                            acc = fmaf(Wr[half + k], d1, acc);
                            acc = fmaf(Wr[2 * half + k], d2, acc);
                            const float gv = (e & 1) ? __high2float(gh[e >> 1]) : __low2float(gh[e >> 1]);
                            v[e] = gv > 0.0f ? acc : 0.0f;
                        }
                        // appearance-embedding gradient, step 1: per-image sums of dZ_dira rows (fp32, unscaled); the lanes of a warp
                        // are consecutive slots, i.e. mostly samples of one ray = one image id
                        if (A.emb_sum) {
                            unsigned todo = __ballot_sync(0xffffffffu, row >= 0);
                            while (todo) {
                                const int leader = __ffs(todo) - 1;
                                const int cur = __shfl_sync(0xffffffffu, id, leader);
                                const bool mine = row >= 0 && id == cur;
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    float t = mine ? v[e] : 0.0f;
#pragma unroll
                                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                                    if (lane == leader) atomicAdd(A.emb_sum + ((size_t)sub0 * A.m.nd.app_count + cur) * half + k0 + e, t);
                                }
                                todo &= ~__ballot_sync(0xffffffffu, mine);
                            }
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) pk[e] = pack_h2(v[2 * e] * S, v[2 * e + 1] * S);
                        const uint4 outv = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(Hsl + (size_t)(k0 >> 3) * (kTileM * 16) + (size_t)r * 16) = outv;
                        *reinterpret_cast<uint4*>(dimg + (size_t)(k0 >> 3) * (kTileM * 16) + (size_t)r * 16) = outv;
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&epi_done[sl]);
                }
            }
            for (int gi = 0; gi < n_gemm; ++gi) {
                const TcGemm& g = P.g[gi];
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl == 1 && !valid1) continue;
                    if (sl == 0) { mbar_wait(&acc_full[0], aph0); aph0 ^= 1; }
                    else         { mbar_wait(&acc_full[1], aph1); aph1 ^= 1; }
                    tc_fence_after();
                    if (warp == 0 && lane == 0) trace_ev(A.desc_swap, 1, 3, sl, gi);   // epilogue: accumulator ready
                    const uint32_t t_acc = t_lane + (uint32_t)sl * 256u;
                    const float* bias = F32 + g.bias_off;
                    const int64_t row = row_[sl], slot = slot_[sl];
                    unsigned char* Hsl = Hs + (size_t)sl * h_bytes;
                    if (kMode == PP_DGRAD) {
                        // dH = accumulator (scaled by S) [+ S dsigma x w_sigma] -> ReLU mask from the activation tape -> fp16 ->
                        // next A operand + gradient tape.  g.bias_off holds the image index (the mask image and the target image
                        // coincide: dZ_l = dH_l where H_l > 0).
                        const int img = g.bias_off;
                        const size_t ioff = (size_t)(t0 + sl) * A.act_tile_bytes + (size_t)img * L * kTileM * 2;
                        const unsigned char* mimg = A.tape_act + ioff;
                        unsigned char* dimg = A.tape_dz + ioff;
                        const float dss = dsig_[sl];
                        const int nslab = (g.n + 63) >> 6;
                        for (int j = 0; j < nslab; ++j) {
                            const int c0 = 64 * j + 16 * part;
                            if (c0 >= g.n) continue;
                            uint32_t v[16];
                            tmem_ld16(t_acc + (uint32_t)c0, v);
                            const size_t po = (size_t)(c0 >> 3) * (kTileM * 16) + (size_t)r * 16;
                            uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
                            if (g.epi != EPI_D_LINEAR) {
                                m0 = *reinterpret_cast<const uint4*>(mimg + po);
                                m1 = *reinterpret_cast<const uint4*>(mimg + po + kTileM * 16);
                            }
                            tmem_ld_wait();
                            float f[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                            if (g.epi == EPI_D_MASK_SIGMA) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) f[i] = fmaf(dss, F32[c0 + i], f[i]);      // F32[0..L) = sigma weights
                            }
                            if (g.epi != EPI_D_LINEAR) {
                                const __half2* h0 = reinterpret_cast<const __half2*>(&m0);
                                const __half2* h1 = reinterpret_cast<const __half2*>(&m1);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float hv = (i & 1) ? __high2float(h0[i >> 1]) : __low2float(h0[i >> 1]);
                                    const float hw = (i & 1) ? __high2float(h1[i >> 1]) : __low2float(h1[i >> 1]);
                                    if (!(hv > 0.0f)) f[i] = 0.0f;
                                    if (!(hw > 0.0f)) f[8 + i] = 0.0f;
                                }
                            }
                            uint32_t pk[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) pk[e] = pack_h2(f[2 * e], f[2 * e + 1]);
                            const uint4 o0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), o1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            *reinterpret_cast<uint4*>(Hsl + po) = o0;
                            *reinterpret_cast<uint4*>(Hsl + po + kTileM * 16) = o1;
                            *reinterpret_cast<uint4*>(dimg + po) = o0;
                            *reinterpret_cast<uint4*>(dimg + po + kTileM * 16) = o1;
                        }
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        // the last GEMM's epilogue is followed by this slot's next head stage (same warps): nobody waits for it
                        if (lane == 0 && gi + 1 < n_gemm) mbar_arrive(&epi_done[sl]);
                        continue;
                    }
                    // training forward: tape image of this GEMM's output (trunk layer gi; then F, then G)
                    unsigned char* timg = nullptr;
                    if (kMode == PP_TRAIN_FWD && g.epi != EPI_RGB)
                        timg = A.tape_act + (size_t)(t0 + sl) * A.act_tile_bytes + (size_t)gi * L * kTileM * 2;
                    if (g.epi == EPI_RGB) {
                        if (part == 0) {
                            uint32_t v[32];
                            tmem_ld32(t_acc, v);
                            tmem_ld_wait();
                            float* tr = kMode == PP_TRAIN_FWD ? A.tape_f32 + (size_t)(t0 + sl) * 5 * kTileM + kTileM + r : nullptr;
                            if (row >= 0) tc_emit_rgb(A.m, A.m.nd.affine ? sub_of(t0 + sl) : 0, row, slot, v, RGBB, sigma_[sl], tr);
                            else if (tr) { tr[0] = 0.5f; tr[kTileM] = 0.5f; tr[2 * kTileM] = 0.5f; }
                        }
                    } else {
                        const bool want_sigma = g.epi == EPI_RELU_SIGMA;
                        const bool publish = !(want_sigma && A.m.sigma_only);
                        const float* sw = SW;
                        float sacc = 0.0f;
                        const int nslab = (g.n + 63) >> 6;
                        for (int j = 0; j < nslab; ++j) {
                            const int c0 = 64 * j + 16 * part;
                            if (c0 < g.n) {
                                const size_t po = (size_t)(c0 >> 3) * (kTileM * 16) + (size_t)r * 16;
                                unsigned char* dst = Hsl + po;
                                unsigned char* gd = timg ? timg + po : nullptr;
                                if (g.epi == EPI_RELU)
                                    epi_piece16<false, true, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, 0, true, gd);
                                else if (g.epi == EPI_LINEAR)
                                    epi_piece16<false, false, false>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, 0, true, gd);
                                else
                                    sacc += epi_piece16<false, true, true>(t_acc + (uint32_t)c0, bias + c0, sw + c0, dst, 0, publish, gd);
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
                                if (kMode == PP_TRAIN_FWD) {
                                    float* tf = A.tape_f32 + (size_t)(t0 + sl) * 5 * kTileM + r;
                                    tf[0] = s;                                                  // pre-activation (with the density noise)
                                    tf[4 * kTileM] = (row >= 0 && A.m.nd.app > 0) ? A.m.src.index(row) : 0.0f;   // image id of the row
                                }
                                if (A.m.sigma_only && row >= 0) {
                                    const int64_t o = (A.m.scatter ? row : slot) * A.m.out_cols;
                                    A.m.out[o] = A.m.slot_w ? sg * A.m.slot_w[slot] : sg;
                                }
                            }
                            asm volatile("bar.sync 1, 512;" ::: "memory");   // SIGP is reused by the other tile slot
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (warp == 0 && lane == 0) trace_ev(A.desc_swap, 1, 4, sl, gi);   // epilogue: this warp done
                    if (lane == 0) mbar_arrive(&epi_done[sl]);
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

#include "mn_mlp_tp.cuh"
#include "mn_mlp_wide.cuh"
#include "mn_train_tc.cuh"

}  // namespace

// =================================================================================================
size_t mn_mlp_tc_workspace(const mn_model* m, int64_t n_tiles128, int precision) {
    TcPlan P;
    if (!build_plan(m->nd, &P)) return 0;
    const size_t planes = precision == MN_PREC_TC_F16X3 ? 2 : 1;
    return mn_align((size_t)n_tiles128 * P.x_tile_bytes * planes, 1024) + 1024;
}

int mn_mlp_tp_program(const NetDims& nd, unsigned int* table_out, int cap_entries, int* info8) {
    TcPlan P;
    if (!build_plan(nd, &P) || nd.L > 256) return MN_ERR_UNSUPPORTED;
    std::vector<uint4> table;
    int counts[4];
    if (!tp_build_program(P, &table, counts)) return MN_ERR_UNSUPPORTED;
    size_t tp_bytes = 0;
    for (int gi = 0; gi < P.n_gemm; ++gi) tp_bytes += (size_t)tp_gemm_bytes(P.g[gi]);
    const TPLayout TL = tp_layout(P);
    const int info[8] = {counts[0], counts[1], counts[2], counts[3], (int)tp_bytes, TL.stages, TL.total, P.x_tile_bytes};
    for (int i = 0; i < 8; ++i) info8[i] = info[i];
    int n = 0;
    for (int i = 0; i < counts[0] && n < cap_entries; ++i, ++n) {
        const uint4 e = table[i];
        table_out[4 * n] = e.x; table_out[4 * n + 1] = e.y; table_out[4 * n + 2] = e.z; table_out[4 * n + 3] = e.w;
    }
    for (int i = 0; i < counts[2] && n < cap_entries; ++i, ++n) {
        const uint4 e = table[kTPMaxProg + i];
        table_out[4 * n] = e.x; table_out[4 * n + 1] = e.y; table_out[4 * n + 2] = e.z; table_out[4 * n + 3] = e.w;
    }
    return MN_OK;
}

int mn_mlp_tc_pack(mn_ctx* ctx, mn_model* m, int sub, cudaStream_t st) {
    TcPlan P;
    if (!build_plan(m->nd, &P)) {
        m->tc_ready = 0;
        return MN_OK;  // configuration only served by the fp32 kernel
    }
    const NetDims& nd = m->nd;
    // per sub-module: [hi plane][lo plane][fp32 block, 256-aligned]
    // 512-wide network: only the wide kernel runs it; its one image ([N half of 256][K/8][256][8]) lives in the hi plane
    const bool wide = nd.L > 256;
    const size_t sub_bytes = mn_align((size_t)P.plane_bytes * 2 + (size_t)(((P.f32_floats * 4 + 255) / 256) * 256), 256);
    if (!m->tc_packed) {
        MN_CUDA(ctx, cudaMalloc(&m->tc_packed, sub_bytes * m->d.n_sub));
        MN_CUDA(ctx, cudaMemsetAsync(m->tc_packed, 0, sub_bytes * m->d.n_sub, st));
        m->tc_sub_bytes = sub_bytes;
    }
    unsigned char* base = (unsigned char*)m->tc_packed + (size_t)sub * sub_bytes;
    const float* Pk = m->packed + (size_t)sub * m->lay.total;
    // TMEM ping-pong kernel (layer_dim <= 256): its own weight plane and, once per model, the role tables
    unsigned char* tp_base = nullptr;
    size_t tp_woff = 0;
    if (!wide) {
        size_t tp_bytes = 0;
        for (int gi2 = 0; gi2 < P.n_gemm; ++gi2) tp_bytes += (size_t)tp_gemm_bytes(P.g[gi2]);
        tp_bytes = mn_align(tp_bytes, 256);
        if (!m->tc_tp) {
            std::vector<uint4> table;
            if (tp_build_program(P, &table, m->tp_n)) {
                MN_CUDA(ctx, cudaMalloc(&m->tc_tp, tp_bytes * m->d.n_sub));
                MN_CUDA(ctx, cudaMemsetAsync(m->tc_tp, 0, tp_bytes * m->d.n_sub, st));
                MN_CUDA(ctx, cudaMalloc(&m->tp_prog, table.size() * sizeof(uint4)));
                MN_CUDA(ctx, cudaMemcpyAsync(m->tp_prog, table.data(), table.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
                MN_CUDA(ctx, cudaStreamSynchronize(st));      // `table` is pageable host memory
                m->tc_tp_sub_bytes = tp_bytes;
            }
        }
        if (m->tc_tp) tp_base = (unsigned char*)m->tc_tp + (size_t)sub * m->tc_tp_sub_bytes;
    }
    float* f32 = reinterpret_cast<float*>(base + (size_t)P.plane_bytes * 2);
    auto pack = [&](const TcGemm& g, const float* wt, int n_src, int k_src, int k_real0, int k_pad0, const float* bias,
                    int n_bias) -> int {
        const int K = g.k[0] + (g.nseg > 1 ? g.k[1] : 0);
        const int64_t n = (int64_t)g.n * K;
        __half* hi = reinterpret_cast<__half*>(base + g.w_off);
        __half* lo = reinterpret_cast<__half*>(base + P.plane_bytes + g.w_off);
        // queued: all images of the sub-module are written by ONE launch (mn_pack_flush in mn_model_set_weights)
        if (wide) {
            mn_pack_push(ctx, PackOp{wt, hi, nullptr, (long long)n, PK_TC_HALF, {n_src, k_src, g.n, K, k_real0, k_pad0, g.n < 256 ? g.n : 256}});
            mn_pack_push(ctx, PackOp{bias, f32 + g.bias_off, nullptr, (long long)P.bstride, PK_TC_F32, {n_bias, 0, 0, 0, 0, 0, 0}});
            return MN_OK;
        }
        mn_pack_push(ctx, PackOp{wt, hi, lo, (long long)n, PK_TC_IMAGE, {n_src, k_src, g.n, K, k_real0, k_pad0, 0}});
        mn_pack_push(ctx, PackOp{bias, f32 + g.bias_off, nullptr, 256, PK_TC_F32, {n_bias, 0, 0, 0, 0, 0, 0}});
        if (tp_base) {       // half-major image of the TMEM ping-pong kernel: [N-half][K/8][nw][8]
            const int nw = g.n < 128 ? g.n : 128, nh = (g.n + 127) / 128;
            mn_pack_push(ctx, PackOp{wt, tp_base + tp_woff, nullptr, (long long)nh * nw * K, PK_TC_HALF, {n_src, k_src, g.n, K, k_real0, k_pad0, nw}});
            tp_woff += tp_gemm_bytes(g);
        }
        return MN_OK;
    };
    int rc, gi = 0;
    for (int i = 0; i < nd.layers; ++i, ++gi) {
        const bool has_pe = (i == 0) || ((nd.skip_mask >> i) & 1);
        if ((rc = pack(P.g[gi], Pk + m->lay.w[i], nd.L, m->lay.kin[i], has_pe ? nd.in_xyz : 0, has_pe ? P.kpe : 0,
                       Pk + m->lay.b[i], nd.L)))
            return rc;
    }
    if (nd.has_dir_a) {
        if ((rc = pack(P.g[gi++], Pk + m->lay.final_w, nd.L, nd.L, 0, 0, Pk + m->lay.final_b, nd.L))) return rc;
        if ((rc = pack(P.g[gi++], Pk + m->lay.dira_w, nd.L / 2, nd.L + nd.aux, 0, 0, Pk + m->lay.dira_b, nd.L / 2))) return rc;
    }
    if ((rc = pack(P.g[gi++], Pk + m->lay.rgb_w, nd.rgb_dim, nd.rgb_in, 0, 0, Pk + m->lay.rgb_b, nd.rgb_dim))) return rc;
    // sigma_w [L] + sigma_b
    mn_pack_push(ctx, PackOp{Pk + m->lay.sigma_w, f32 + P.sigma_w_off, nullptr, (long long)nd.L, PK_TC_F32, {nd.L, 0, 0, 0, 0, 0, 0}});
    mn_pack_push(ctx, PackOp{Pk + m->lay.sigma_b, f32 + P.sigma_w_off + nd.L, nullptr, 4, PK_TC_F32, {1, 0, 0, 0, 0, 0, 0}});
    m->tc_ready = 1;

    // ---- data-gradient images of the tensor-core training path (transposed weights, single fp16 plane + fp32 block)
    TcPlan D;
    m->train_tc_ok = 0;
    if (!wide && build_dgrad_plan(nd, &D)) {
        if (!m->tc_dgrad) {
            MN_CUDA(ctx, cudaMalloc(&m->tc_dgrad, (size_t)D.sub_bytes * m->d.n_sub));
            MN_CUDA(ctx, cudaMemsetAsync(m->tc_dgrad, 0, (size_t)D.sub_bytes * m->d.n_sub, st));
            m->tc_dgrad_sub_bytes = (size_t)D.sub_bytes;
        }
        unsigned char* db = (unsigned char*)m->tc_dgrad + (size_t)sub * D.sub_bytes;
        const float* Q = m->packed_bwd + (size_t)sub * m->blay.total;
        auto packd = [&](const TcGemm& g, const float* wd, int ld) -> int {
            mn_pack_push(ctx, PackOp{wd, db + g.w_off, nullptr, (long long)g.n * g.k[0], PK_DGRAD, {ld, g.n, g.k[0], 0, 0, 0, 0}});
            return MN_OK;
        };
        int di = 0;
        if ((rc = packd(D.g[di++], Q + m->blay.dira_f, nd.L))) return rc;         // [L/2][L]
        if ((rc = packd(D.g[di++], Q + m->blay.final_w, nd.L))) return rc;        // [L][L]
        for (int l = nd.layers - 1; l >= 1; --l)
            if ((rc = packd(D.g[di++], Q + m->blay.w[l], nd.L))) return rc;       // [L][L]: hidden-part columns of layer l
        float* df32 = reinterpret_cast<float*>(db + D.f32_off);
        mn_pack_push(ctx, PackOp{Pk + m->lay.sigma_w, df32, nullptr, (long long)nd.L, PK_TC_F32, {nd.L, 0, 0, 0, 0, 0, 0}});
        mn_pack_push(ctx, PackOp{Pk + m->lay.rgb_w, df32 + nd.L, nullptr, (long long)3 * (nd.L / 2), PK_RGBW, {nd.L / 2, 3, 0, 0, 0, 0, 0}});
        m->train_tc_ok = 1;
    }
    return MN_OK;
}

int mn_mlp_tc_launch(mn_ctx* ctx, mn_model* m, const MlpArgs& a, int64_t n_tiles128, int precision, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
    TcArgs A{};
    if (!build_plan(a.nd, &A.plan) || !m->tc_ready)
        return mn_fail(ctx, MN_ERR_UNSUPPORTED,
                       "tensor-core MLP path covers layer_dim 64..256 (multiple of 64) or 512 and rgb_dim <= 32; "
                       "use precision 'fp32' for this model");
    if (a.nd.L > 256 && precision == MN_PREC_TC_F16X3)
        return mn_fail(ctx, MN_ERR_UNSUPPORTED,
                       "precision 'tc_f16x3' covers layer_dim <= 256; use 'tc_f16' or 'fp32' for the 512-wide network");
    if (n_tiles128 <= 0) return MN_OK;
    TcPlan& P = A.plan;
    const int split = precision == MN_PREC_TC_F16X3 ? 1 : 0;
    P.sub_bytes = (int)m->tc_sub_bytes;
    // the packed layout always holds both planes; tell the kernel where the fp32 block is
    A.m = a;
    A.wpack = (const unsigned char*)m->tc_packed;
    A.split = split;
    static int desc_swap = -1;
    if (desc_swap < 0) {
        const char* e = getenv("MN_TC_TRACE");
        desc_swap = e ? atoi(e) : 0;      // bit 0: in-kernel timeline + SM-clock stamp
    }
    A.desc_swap = desc_swap;
    A.n_tiles_cap = n_tiles128;
    const size_t need = mn_mlp_tc_workspace(m, n_tiles128, precision);
    if (ws_bytes < need || !ws) return mn_fail(ctx, MN_ERR_WORKSPACE, "mn_mlp_tc_launch: workspace too small");
    uintptr_t wp = ((uintptr_t)ws + 1023) / 1024 * 1024;
    __half* ximg = reinterpret_cast<__half*>(wp);
    A.ximg = ximg;
    A.x_plane_halves = (int64_t)n_tiles128 * (P.kpe + P.kaux) * kTileM;

    // ---- kernel selection: the ping-pong kernel for layer_dim <= 256 (MN_TC_PINGPONG=0: the single-tile kernel, kept as
    // the cross-check of the variants test), the wide kernel for 512, the split kernel for tc_f16x3
    static int use_pp = -1;
    if (use_pp < 0) {
        const char* e = getenv("MN_TC_PINGPONG");
        use_pp = (e && e[0] == '0') ? 0 : 1;
    }
    // The TMEM ping-pong kernel (mn_mlp_tp.cuh) is the default inference kernel for layer_dim <= 256; MN_TC_TP=0 selects the
    // shared-memory ping-pong kernel (which also serves the two training modes) - variants test, A/B runs
    static int use_tp = -1;
    if (use_tp < 0) {
        const char* e = getenv("MN_TC_TP");
        use_tp = (e && e[0] == '0') ? 0 : 1;
    }
    const TPLayout TL = tp_layout(P);
    const bool run_tp = !split && P.L <= 256 && use_tp && use_pp && m->tc_tp && TL.stages >= 8;
    A.tpack = (const unsigned char*)m->tc_tp;
    A.tp_sub_bytes = m->tc_tp_sub_bytes;
    A.tp_prog = (const uint4*)m->tp_prog;
    for (int i = 0; i < 4; ++i) A.tp_n[i] = m->tp_n[i];
    const PPLayout PL = pp_layout(P);
    const bool run_pp = !split && P.L <= 256 && use_pp && PL.total <= kSmemMax && PL.stages >= 3 &&
                        pp_prog_entries(P, P.n_gemm) + 2 <= kPPMaxProg;

    const size_t enc_sm = (size_t)P.x_tile_bytes * (split ? 2 : 1);
    MN_CUDA(ctx, cudaFuncSetAttribute(tc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_sm));
    const NetDims& ndE = a.nd;
    const bool fast_shape = !split && ndE.xyz_dim == 3 && ndE.nf_xyz == 12 && ndE.nf_dir == 4 && ndE.app == 48 && ndE.app_in_dira &&
                            (m->lay.emb % 4) == 0;
    if (fast_shape)
        tc_encode_fast_kernel<3, 12, 4, 48><<<(unsigned)n_tiles128, kTileM, 0, st>>>(a, ximg);
    else
        tc_encode_kernel<<<(unsigned)n_tiles128, kTileM, enc_sm, st>>>(a, P.kpe, P.kaux, split, ximg, A.x_plane_halves);
    MN_LAUNCH_CHECK(ctx);

    if (P.L > 256) {
        const WLayout WL = w_layout(P);
        if (WL.total > kSmemMax) return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core MLP (512-wide): shared-memory budget exceeded");
        // MN_TC_CLUSTER=2: clusters of two CTAs share one multicast weight stream (the router's bucket alignment
        // guarantees that tiles 2p and 2p+1 belong to one sub-module).  Halves the L2 reads but measured 2-3 % SLOWER
        // on B200: every SM still RECEIVES all weight bytes, and the ~32 B/clk/SM delivery rate is what binds these kernels
        //  - multicast saves L2 reads, not deliveries.  The default is one CTA per cluster.
        static int cs = -1;
        if (cs < 0) {
            const char* e = getenv("MN_TC_CLUSTER");
            cs = (e && e[0] == '2') ? 2 : 1;
        }
        mn_prof_begin(ctx, st);
        if (cs == 2) {
            MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WL.total));
            const int64_t groups = (n_tiles128 + 1) / 2;
            const unsigned ncl = (unsigned)(groups < ctx->sm_count / 2 ? groups : ctx->sm_count / 2);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(2 * ncl);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = (size_t)WL.total;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            MN_CUDA(ctx, cudaLaunchKernelEx(&cfg, tc_mlp_wide_kernel<2>, A));
        } else {
            MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WL.total));
            const unsigned grid_w = (unsigned)(n_tiles128 < ctx->sm_count ? n_tiles128 : ctx->sm_count);
            tc_mlp_wide_kernel<1><<<grid_w, kThreads, WL.total, st>>>(A);
        }
        mn_prof_end(ctx, st);
        MN_LAUNCH_CHECK(ctx);
        return MN_OK;
    }
    const SmemLayout SL = smem_layout(P, split != 0);
    const int total = SL.total;
    if (total > kSmemMax || SL.stages < 2) return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core MLP: shared-memory budget exceeded");
    const unsigned grid = (unsigned)(n_tiles128 < ctx->sm_count ? n_tiles128 : ctx->sm_count);
    if (split) {
        MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, total));
        mn_prof_begin(ctx, st);
        tc_mlp_kernel<true><<<grid, kThreads, total, st>>>(A);
    } else {
        if (run_tp) {
            const unsigned grid_tp = (unsigned)((n_tiles128 + 1) / 2 < ctx->sm_count ? (n_tiles128 + 1) / 2 : ctx->sm_count);
            MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_tp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL.total));
            mn_prof_begin(ctx, st);
            tc_mlp_tp_kernel<<<grid_tp, kTPThreads, TL.total, st>>>(A);
        } else if (run_pp) {
            const unsigned grid_pp = (unsigned)((n_tiles128 + 1) / 2 < ctx->sm_count ? (n_tiles128 + 1) / 2 : ctx->sm_count);
            MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_pp_kernel<PP_INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, PL.total));
            mn_prof_begin(ctx, st);
            tc_mlp_pp_kernel<PP_INFER><<<grid_pp, kPPThreads, PL.total, st>>>(A);
        } else {
            MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, total));
            mn_prof_begin(ctx, st);
            tc_mlp_kernel<false><<<grid, kThreads, total, st>>>(A);
        }
    }
    mn_prof_end(ctx, st);
    MN_LAUNCH_CHECK(ctx);
    return MN_OK;
}

// =================================================================================================
// tensor-core training path: host side (kernels in mn_train_tc.cuh)
// =================================================================================================
size_t mn_train_tc_x_tile_bytes(const mn_model* m) {
    TcPlan P;
    return build_plan(m->nd, &P) ? (size_t)P.x_tile_bytes : 0;
}
size_t mn_train_tc_act_tile_bytes(const mn_model* m) {
    const NetDims& nd = m->nd;
    return (size_t)(nd.layers + 1) * nd.L * kTileM * 2 + (size_t)(nd.L / 2) * kTileM * 2;
}

// recording forward: encoder tiles and every layer's activations land in the caller's tape
int mn_mlp_tc_launch_train(mn_ctx* ctx, mn_model* m, const MlpArgs& a, int64_t n_tiles128, const TrainTcTape& tape, cudaStream_t st) {
    TcArgs A{};
    if (!m->train_tc_ok || !build_plan(a.nd, &A.plan) || !m->tc_ready)
        return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core training covers layer_dim 256 with a direction / appearance head and rgb_dim 3; use train precision 'fp32'");
    if (n_tiles128 <= 0) return MN_OK;
    TcPlan& P = A.plan;
    P.sub_bytes = (int)m->tc_sub_bytes;
    A.m = a;
    A.wpack = (const unsigned char*)m->tc_packed;
    A.ximg = reinterpret_cast<const __half*>(tape.xreg);
    A.x_plane_halves = 0;
    A.n_tiles_cap = n_tiles128;
    A.tape_act = tape.act;
    A.tape_f32 = tape.f32;
    A.act_tile_bytes = (int64_t)mn_train_tc_act_tile_bytes(m);
    A.layers = a.nd.layers;
    const size_t enc_sm = (size_t)P.x_tile_bytes;
    MN_CUDA(ctx, cudaFuncSetAttribute(tc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_sm));
    const NetDims& ndE = a.nd;
    const bool fast_shape = ndE.xyz_dim == 3 && ndE.nf_xyz == 12 && ndE.nf_dir == 4 && ndE.app == 48 && ndE.app_in_dira && (m->lay.emb % 4) == 0;
    __half* ximg = reinterpret_cast<__half*>(tape.xreg);
    if (fast_shape) tc_encode_fast_kernel<3, 12, 4, 48><<<(unsigned)n_tiles128, kTileM, 0, st>>>(a, ximg);
    else tc_encode_kernel<<<(unsigned)n_tiles128, kTileM, enc_sm, st>>>(a, P.kpe, P.kaux, 0, ximg, 0);
    MN_LAUNCH_CHECK(ctx);
    const PPLayout PL = pp_layout(P);
    if (PL.total > kSmemMax || PL.stages < 3 || pp_prog_entries(P, P.n_gemm) + 2 > kPPMaxProg)
        return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core training forward: shared-memory budget exceeded");
    const unsigned grid = (unsigned)((n_tiles128 + 1) / 2 < ctx->sm_count ? (n_tiles128 + 1) / 2 : ctx->sm_count);
    MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_pp_kernel<PP_TRAIN_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, PL.total));
    mn_prof_begin(ctx, st);
    tc_mlp_pp_kernel<PP_TRAIN_FWD><<<grid, kPPThreads, PL.total, st>>>(A);
    mn_prof_end(ctx, st);
    MN_LAUNCH_CHECK(ctx);
    return MN_OK;
}

// backward workspace: [gradient records][head gradients fp32 [n_tiles][4][128]][embedding sums][scale]
static size_t train_tc_emb_floats(const mn_model* m) {
    return m->nd.app_in_dira ? (size_t)m->d.n_sub * m->nd.app_count * (m->nd.L / 2) : 0;
}
size_t mn_train_tc_backward_workspace(const mn_model* m, int64_t n_tiles128) {
    return mn_align((size_t)n_tiles128 * mn_train_tc_act_tile_bytes(m)) + mn_align((size_t)n_tiles128 * 4 * kTileM * sizeof(float)) +
           mn_align(train_tc_emb_floats(m) * sizeof(float) + 256) + 1024;
}

int mn_train_tc_backward(mn_ctx* ctx, mn_model* m, const BwdArgs& a, int64_t n_tiles128, const TrainTcTape& tape, void* ws, size_t ws_bytes,
                         cudaStream_t st) {
    TcArgs A{};
    const NetDims& nd = a.nd;
    if (!m->train_tc_ok || !build_dgrad_plan(nd, &A.plan)) return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core backward: unsupported network shape");
    if (n_tiles128 <= 0) return MN_OK;
    if (!ws || ws_bytes < mn_train_tc_backward_workspace(m, n_tiles128)) return mn_fail(ctx, MN_ERR_WORKSPACE, "mn_train_tc_backward: workspace too small");
    const size_t act_tile = mn_train_tc_act_tile_bytes(m);
    char* wp = (char*)(((uintptr_t)ws + 255) / 256 * 256);
    unsigned char* dz = (unsigned char*)wp;               wp += mn_align((size_t)n_tiles128 * act_tile);
    float* gf32 = (float*)wp;                             wp += mn_align((size_t)n_tiles128 * 4 * kTileM * sizeof(float));
    float* emb_sum = (float*)wp;                          wp += mn_align(train_tc_emb_floats(m) * sizeof(float) + 256) - 256;
    float* scale = (float*)wp;
    if (train_tc_emb_floats(m)) MN_CUDA(ctx, cudaMemsetAsync(emb_sum, 0, train_tc_emb_floats(m) * sizeof(float), st));

    mn_prof_begin(ctx, st);   // bench.py --mode train: the whole backward of the MLP stage timed as one span
    // ---- gradient scale (power of two) from the upstream gradient
    tc_grad_scale_kernel<<<1, 1024, 0, st>>>(a.grad_out, a.grad_rows * a.out_cols, scale);
    MN_LAUNCH_CHECK(ctx);

    // ---- data gradients
    TcPlan& D = A.plan;
    A.m = MlpArgs{};
    A.m.nd = nd;
    A.m.slot_row = a.slot_row;
    A.m.slot_w = a.slot_w;
    A.m.counters = a.counters;
    A.m.n_sub = a.n_sub;
    A.m.fixed_sub = a.fixed_sub;
    A.m.B = a.B;
    A.m.out_cols = a.out_cols;
    A.wpack = (const unsigned char*)m->tc_dgrad;
    A.n_tiles_cap = n_tiles128;
    A.tape_act = tape.act;
    A.tape_f32 = tape.f32;
    A.tape_dz = dz;
    A.tape_gf32 = gf32;
    A.grad_out = a.grad_out;
    A.emb_sum = nd.app_in_dira ? emb_sum : nullptr;
    A.scale = scale;
    A.act_tile_bytes = (int64_t)act_tile;
    A.layers = nd.layers;
    const PPLayout PL = pp_layout(D);
    if (PL.total > kSmemMax || PL.stages < 3 || pp_prog_entries(D, D.n_gemm) + 2 > kPPMaxProg)
        return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core backward: shared-memory budget exceeded");
    const unsigned grid = (unsigned)((n_tiles128 + 1) / 2 < ctx->sm_count ? (n_tiles128 + 1) / 2 : ctx->sm_count);
    MN_CUDA(ctx, cudaFuncSetAttribute(tc_mlp_pp_kernel<PP_DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, PL.total));
    tc_mlp_pp_kernel<PP_DGRAD><<<grid, kPPThreads, PL.total, st>>>(A);
    MN_LAUNCH_CHECK(ctx);

    // ---- weight gradients: one item per (Linear input segment, 128-channel output half)
    WgArgs W{};
    const int L = nd.L;
    const int img_bytes = L * kTileM * 2;
    TcPlan F;
    build_plan(nd, &F);
    int ni = 0;
    auto item = [&](int dz_img, int half, int x_region, int x_off, int n, int n_real, int w_off, int k_in, int in0, int b_off) {
        WgItem& it = W.item[ni++];
        it.dz_off = dz_img * img_bytes + half * 16 * (kTileM * 16);
        it.x_region = x_region;
        it.x_off = x_off;
        it.n = n;
        it.n_real = n_real;
        it.w_off = w_off + half * 128 * k_in + in0;
        it.k_in = k_in;
        it.b_off = b_off >= 0 ? b_off + half * 128 : -1;
    };
    for (int i = 0; i < nd.layers; ++i) {
        const bool skip = i > 0 && ((nd.skip_mask >> i) & 1);
        const int k_in = a.lay.kin[i];
        for (int h = 0; h < L / 128; ++h) {
            if (i == 0) item(i, h, 1, 0, F.kpe, nd.in_xyz, a.lay.w[i], k_in, 0, a.lay.b[i]);
            else if (skip) {
                item(i, h, 1, 0, F.kpe, nd.in_xyz, a.lay.w[i], k_in, 0, a.lay.b[i]);                       // cat[PE, h]: PE columns first
                item(i, h, 0, (i - 1) * img_bytes, L, L, a.lay.w[i], k_in, nd.in_xyz, -1);
            } else item(i, h, 0, (i - 1) * img_bytes, L, L, a.lay.w[i], k_in, 0, a.lay.b[i]);
        }
    }
    for (int h = 0; h < L / 128; ++h) item(nd.layers, h, 0, (nd.layers - 1) * img_bytes, L, L, a.lay.final_w, L, 0, a.lay.final_b);
    // dir_a_encoding: L/2 = 128 output channels (one half); input = cat[final (L), dir PE + embedding (aux)]
    item(nd.layers + 1, 0, 0, nd.layers * img_bytes, L, L, a.lay.dira_w, L + nd.aux, 0, a.lay.dira_b);
    item(nd.layers + 1, 0, 1, (F.kpe / 8) * (kTileM * 16), F.kaux, nd.aux, a.lay.dira_w, L + nd.aux, L, -1);
    if (ni > kWgMaxItems) return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core backward: too many weight-gradient items");
    W.n_items = ni;
    W.act = tape.act;
    W.dz = dz;
    W.xreg = tape.xreg;
    W.act_tile_bytes = (int64_t)act_tile;
    W.x_tile_bytes = (int64_t)F.x_tile_bytes;
    // tiles the forward pass really wrote: all bucketed tiles when routed, ceil(rows / 128) otherwise
    const int64_t tiles_used = a.counters ? n_tiles128 : mn_cdiv(a.B, (int64_t)kTileM);
    W.counters = a.counters;
    W.n_tiles = tiles_used;
    W.fixed_sub = a.fixed_sub;
    W.gw = a.gw;
    W.sub_stride = a.lay.total;
    W.scale = scale;
    // chunks: enough CTAs to fill the machine about three times over (each streams its tiles once; results are fp32 atomics)
    const int n_sub = a.counters ? a.n_sub : 1;
    int64_t chunks = mn_cdiv((int64_t)ctx->sm_count * 3, (int64_t)ni * n_sub);
    if (chunks < 1) chunks = 1;
    int64_t chunk_tiles = mn_cdiv(mn_cdiv(tiles_used, n_sub), chunks);
    if (chunk_tiles < 8) chunk_tiles = 8;
    W.chunk_tiles = (int)chunk_tiles;
    const unsigned gx = (unsigned)mn_cdiv(tiles_used, chunk_tiles);
    const int wg_smem = 2 * kWgStageBytes + 6144 + 256;
    MN_CUDA(ctx, cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem));
    tc_wgrad_kernel<<<dim3(gx, (unsigned)ni, (unsigned)n_sub), kWgThreads, wg_smem, st>>>(W);
    MN_LAUNCH_CHECK(ctx);

    // ---- sigma / rgb heads and the appearance embedding
    HeadsArgs H{};
    H.act = tape.act;
    H.gf32 = gf32;
    H.act_tile_bytes = (int64_t)act_tile;
    H.L = L;
    H.layers = nd.layers;
    H.counters = a.counters;
    H.n_tiles = tiles_used;
    H.fixed_sub = a.fixed_sub;
    H.chunk_tiles = 16;
    H.gw = a.gw;
    H.sub_stride = a.lay.total;
    H.sigma_w = a.lay.sigma_w; H.sigma_b = a.lay.sigma_b; H.rgb_w = a.lay.rgb_w; H.rgb_b = a.lay.rgb_b;
    tc_heads_wgrad_kernel<<<dim3((unsigned)mn_cdiv(tiles_used, 16), (unsigned)n_sub), 256, 0, st>>>(H);
    MN_LAUNCH_CHECK(ctx);
    if (nd.app_in_dira) {
        tc_emb_grad_kernel<<<dim3((unsigned)nd.app_count, (unsigned)a.n_sub), 64, 0, st>>>(emb_sum, a.packed_bwd, a.blay.total, a.blay.dira_e, L / 2, nd.app,
                                                                                          nd.app_count, a.gw, a.lay.total, a.lay.emb);
        MN_LAUNCH_CHECK(ctx);
    }
    mn_prof_end(ctx, st);
    return MN_OK;
}

extern "C" int mn_debug_read_clock(unsigned long long* out4) {
    cudaDeviceSynchronize();
    if (out4) cudaMemcpyFromSymbol(out4, g_clk, sizeof(unsigned long long) * 4);
    return 0;
}

extern "C" int mn_debug_read_trace(unsigned long long* out, unsigned int* counts, int reset) {
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 4 * 4096);
    if (counts) cudaMemcpyFromSymbol(counts, g_trace_n, sizeof(unsigned int) * 2);
    if (reset) {
        unsigned int z[2] = {0, 0};
        cudaMemcpyToSymbol(g_trace_n, z, sizeof(z));
    }
    return 0;
}

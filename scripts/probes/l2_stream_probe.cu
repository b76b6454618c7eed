// Micro-probe: how fast can one SM pull an L2-resident 1.2 MB weight image into shared memory, over and over, while all
// 148 SMs do the same?  (a) 1-D TMA bulk copies (cp.async.bulk) through an mbarrier ring, (b) cp.async 16 B per lane
// (LDGSTS) by N loader warps, (c) both at once.  Prints bytes / clk / SM for several ring depths and copy sizes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream_probe l2_stream_probe.cu && ./l2_stream_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode bit 0: TMA ring, bit 1: cp.async loader warps.  Each "consumer" just waits for a stage and releases it.
// grid = 148 CTAs x 256 threads: warp 0 lane 0 = TMA producer, warp 1 = consumer of the TMA ring, warps 2..2+nlw-1 = cp.async loaders
// (each loader warp runs its own private double-buffered stream - pure bandwidth).
__global__ void __launch_bounds__(256, 1) probe(const unsigned char* __restrict__ w, size_t image_bytes, int stage_bytes, int stages, int n_copies,
                                                int mode, int nlw, int lw_iters, unsigned long long* clk_out, int nprod, int split2) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;
    unsigned char* ring = smem + 1024;
    unsigned char* lbuf = ring + (size_t)stages * stage_bytes;          // loader warps' buffers: nlw x 2 x 4096 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    const unsigned char* base = w + (size_t)(blockIdx.x % 8) * image_bytes;      // 8 different images, like 8 sub-modules
    if ((mode & 1) && (warp == 0 || (warp >= 4 && warp < 4 + nprod - 1)) && lane == 0) {
        // producer p of nprod handles copies i with i % nprod == p (each walks the whole schedule, acts on its own stages)
        const int p = warp == 0 ? 0 : warp - 3;
        int st = 0; uint32_t ph = 0; size_t off = 0;
        for (int i = 0; i < n_copies; ++i) {
            if (i % nprod == p) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], (uint32_t)stage_bytes);
                if (split2) {          // two half-size copies per stage (like weights + features)
                    bulk_g2s(ring + (size_t)st * stage_bytes, base + off, (uint32_t)stage_bytes / 2, &full[st]);
                    bulk_g2s(ring + (size_t)st * stage_bytes + stage_bytes / 2, base + off + stage_bytes / 2, (uint32_t)stage_bytes / 2, &full[st]);
                } else {
                    bulk_g2s(ring + (size_t)st * stage_bytes, base + off, (uint32_t)stage_bytes, &full[st]);
                }
            }
            off += stage_bytes; if (off + stage_bytes > image_bytes) off = 0;
            if (++st == stages) { st = 0; ph ^= 1; }
        }
    } else if ((mode & 1) && warp == 1 && lane == 0) {
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < n_copies; ++i) {
            mbar_wait(&full[st], ph);
            mbar_arrive(&empty[st]);
            if (++st == stages) { st = 0; ph ^= 1; }
        }
    } else if ((mode & 2) && warp >= 2 && warp < 2 + nlw) {
        // each loader warp: lw_iters groups of 8 x 512 B (cp.async 16 B per lane), 2 groups in flight
        unsigned char* mybuf = lbuf + (size_t)(warp - 2) * 8192;
        size_t off = (size_t)(warp - 2) * 65536 % image_bytes;
        for (int it = 0; it < lw_iters; ++it) {
            unsigned char* dst = mybuf + (it & 1) * 4096;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + j * 512 + lane * 16)), "l"(base + off + j * 512 + lane * 16) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            off += 4096; if (off + 4096 > image_bytes) off = 0;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

int main() {
    const size_t image = 1228800;            // ~1.2 MB fp16 weights of one sub-module
    unsigned char* w; cudaMalloc(&w, image * 8); cudaMemset(w, 1, image * 8);
    unsigned long long* clk; cudaMalloc(&clk, 148 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int mode, stage_bytes, stages, nlw, nprod, split2; };
    const Cfg cfgs[] = {{1, 16384, 5, 0, 1, 0}, {1, 16384, 5, 0, 2, 0}, {1, 16384, 5, 0, 3, 0}, {1, 16384, 5, 0, 4, 0},
                        {1, 16384, 5, 0, 1, 1}, {1, 16384, 5, 0, 2, 1},
                        {1, 8192, 10, 0, 1, 0}, {1, 8192, 10, 0, 2, 0}, {1, 8192, 10, 0, 4, 0},
                        {1, 4096, 16, 0, 4, 0}, {1, 32768, 4, 0, 1, 0}, {1, 32768, 4, 0, 2, 0},
                        {2, 16384, 1, 4, 1, 0}, {3, 16384, 5, 2, 1, 0}, {3, 16384, 5, 2, 2, 0}};
    for (const Cfg& c : cfgs) {
        const size_t tma_bytes = (size_t)64 << 20;                      // per CTA
        const int n_copies = (c.mode & 1) ? (int)(tma_bytes / c.stage_bytes) : 0;
        const int lw_iters = (c.mode & 2) ? (int)(((size_t)16 << 20) / 4096) : 0;    // 16 MB per loader warp
        const size_t sm = 1024 + (size_t)c.stages * c.stage_bytes + 8 * 8192;
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<148, 256, sm>>>(w, image, c.stage_bytes, c.stages, n_copies, c.mode, c.nlw, lw_iters, clk, c.nprod, c.split2);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0, mx = 0; for (int i = 0; i < 148; ++i) { avg += h[i]; if (h[i] > mx) mx = (double)h[i]; }
        avg /= 148;
        const double bytes = (double)n_copies * c.stage_bytes + (double)c.nlw * lw_iters * 4096.0;
        printf("mode %d (1=TMA 2=cp.async 3=both) stage %5d B x %2d stages, %d TMA producer warps%s, %d loader warps: %.1f B/clk/SM (avg), %.1f (slowest SM); TMA part %.1f MB, cp.async part %.1f MB per SM\n",
               c.mode, c.stage_bytes, c.stages, c.nprod, c.split2 ? " (2 copies per stage)" : "", c.nlw, bytes / avg, bytes / mx, n_copies * (double)c.stage_bytes / 1e6, c.nlw * lw_iters * 4096.0 / 1e6);
    }
    return 0;
}

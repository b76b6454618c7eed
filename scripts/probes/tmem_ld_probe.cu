// Probe: how fast can the epilogue warps drain tensor memory?  tcgen05.ld.32x32b.x16 / .x32 issued by W warps of one CTA per SM
// (warp w reads its own 32-lane quarter), `depth` loads in flight between two tcgen05.wait::ld, optionally with a stream of
// tcgen05.mma (A from TMEM or from shared memory) running on the same SM.  Prints bytes / clk / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_probe tmem_ld_probe.cu && ./tmem_ld_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}

// mode: 0 = x16 one load per wait, 1 = x16 two loads per wait, 2 = x32 one load per wait
// mma: 0 = none, 1 = SS-mode N=128 stream into columns [256,384), 2 = TS-mode (A from TMEM columns [384,512))
__global__ void __launch_bounds__(576, 1) probe(int mode, int mma, int iters, unsigned long long* out, unsigned* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(smem + 65536)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(smem + 65536 + 8)));
        *reinterpret_cast<volatile int*>(smem + 65536 + 64) = 0;
        *reinterpret_cast<volatile int*>(smem + 65536 + 68) = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    const long long t0 = clock64();
    unsigned acc = 0;
    if (warp < 16) {
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);   // 32-column piece of [0,128)
        for (int it = 0; it < iters; ++it) {
            uint32_t v[32];
            if (mode == 0) {
                ld16(tl, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] ^ v[15];
                ld16(tl + 16, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] ^ v[15];
            } else if (mode == 1) {
                ld16(tl, v); ld16(tl + 16, v + 16); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] ^ v[31];
            } else {
                ld32(tl, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] ^ v[31];
            }
        }
    } else if (warp == 17 && mma) {
        // a continuous stream of N = 128, K = 16 MMAs (64 clk each at peak): 8 per loader iteration is more than the loaders need
        const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(smem), 2048, 128), bd = make_desc(smem_u32(smem) + 32768, 2048, 128);
        if (lane == 0) {
            // batches of 16 MMAs, each committed to one of two mbarriers; wait for the batch before the previous one, so that
            // 16..32 MMAs are always queued.  Runs until all loader warps are done.
            uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
            uint32_t ph[2] = {0, 0};
            int batch = 0;
            while (*reinterpret_cast<volatile int*>(smem + 65536 + 64) < 16) {
                const int b = batch & 1;
                if (batch >= 2) {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[b])), "r"(ph[b]) : "memory");
                    ph[b] ^= 1;
                }
                for (int i = 0; i < 16; ++i) {
                    if (mma == 1)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem + 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(tmem + 256), "r"(tmem + 384), "l"(bd), "r"(idesc), "r"(1) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[b])) : "memory");
                ++batch;
            }
            for (int k = 0; k < 2 && k < batch; ++k) {       // drain the last two batches
                const int b = (batch - 1 - k) & 1;
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[b])), "r"(ph[b]) : "memory");
                ph[b] ^= 1;
            }
            *reinterpret_cast<volatile int*>(smem + 65536 + 68) = batch;
        }
    }
    const long long t1 = clock64();
    if (warp < 16 && lane == 0) {
        out[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
        atomicAdd(reinterpret_cast<int*>(smem + 65536 + 64), 1);
    }
    if (acc == 0x12345u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) {
        if (lane == 0 && blockIdx.x == 0) out[148 * 16] = (unsigned long long)*reinterpret_cast<volatile int*>(smem + 65536 + 68);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

int main() {
    unsigned long long* out;
    unsigned* sink;
    cudaMalloc(&out, (148 * 16 + 1) * 8);
    cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    const int iters = 2000;
    const char* mn[3] = {"x16, wait after each", "x16 x2, one wait", "x32, one wait"};
    const char* mm[3] = {"no MMA", "SS-mode MMA stream", "TS-mode MMA stream (A from TMEM)"};
    for (int mma = 0; mma < 3; ++mma)
        for (int mode = 0; mode < 3; ++mode) {
            probe<<<148, 576, 66 * 1024>>>(mode, mma, iters, out, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            unsigned long long h[148 * 16];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double mx = 0;
            for (int i = 0; i < 148 * 16; ++i) mx = h[i] > mx ? (double)h[i] : mx;
            // per iteration every warp reads 32 columns x 32 lanes x 4 B = 4 KiB; 16 warps -> 64 KiB per SM
            unsigned long long nb = 0;
            cudaMemcpy(&nb, out + 148 * 16, 8, cudaMemcpyDeviceToHost);
            printf("%-22s | %-34s : %6.1f B/clk/SM  (%.0f clk per 64 KiB drain)", mn[mode], mm[mma], 65536.0 * iters / mx, mx / iters);
            if (mma) printf("   MMA stream: %.1f clk per N=128 K=16 MMA (64 = peak)", mx / (16.0 * (double)nb));
            printf("\n");
        }
    return 0;
}

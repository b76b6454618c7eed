// Probe: does a chain of tcgen05.mma that accumulate into the SAME tensor-memory tile run at the tensor pipe's throughput, or
// at its latency?  One thread per SM issues `iters` MMAs (M = 128, K = 16, N = 128 or 256; A from shared memory or from tensor
// memory) round-robin over `nacc` independent accumulators; nothing else runs on the SM.  Prints clk per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_chain_probe mma_chain_probe.cu && ./mma_chain_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(bd), "r"(idesc) : "memory");
}

// ts: 0 = A from shared memory, 1 = A from tensor memory.  N = 128: accumulators at columns 0,128,256 (A at 384); N = 256: 0,256.
template <int N, int NACC, int TS>
__global__ void __launch_bounds__(128, 1) probe(int iters, unsigned long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(smem), 2048, 128), bd = make_desc(smem_u32(smem) + 32768, N * 16, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += 4 * NACC) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    if (TS) mma_ts(tmem + (uint32_t)(a * N), tmem + 384u + (uint32_t)(8 * u), bd + (uint64_t)(u * 2 * N), idesc);
                    else mma_ss(tmem + (uint32_t)(a * N), ad + (uint64_t)(u * 256), bd + (uint64_t)(u * 2 * N), idesc);
                }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N, int NACC, int TS>
void run(unsigned long long* out) {
    const int iters = 24000;     // multiple of 4 * NACC for NACC in {1, 2, 3}
    cudaFuncSetAttribute(probe<N, NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    probe<N, NACC, TS><<<148, 128, 96 * 1024>>>(iters, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    unsigned long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? (double)h[i] : mx;
    printf("N = %3d, A from %-6s, %d accumulator(s) round-robin: %6.1f clk per MMA  (throughput floor %d)\n", N, TS ? "TMEM" : "smem", NACC,
           mx / iters, N / 2);
}

int main() {
    unsigned long long* out;
    cudaMalloc(&out, 148 * 8);
    run<128, 1, 0>(out); run<128, 2, 0>(out); run<128, 3, 0>(out);
    run<128, 1, 1>(out); run<128, 2, 1>(out); run<128, 3, 1>(out);
    run<256, 1, 0>(out); run<256, 2, 0>(out);
    return 0;
}

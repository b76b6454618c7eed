// Probe: tcgen05.mma with MN-major (transposed) A and B taken from the row-major "tile image" layout the MLP kernels use
// ([cols/8][128 rows][8] fp16: element (row, col) at (col/8)*2048 + row*16 + (col%8)*2).  Computes D[m][n] = sum_rows A[row][m] * B[row][n]
// (M = 128 columns of A, N columns of B, K = 128 rows) - the weight-gradient contraction dW = dZ^T X - and checks it on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mn_major_probe mn_major_probe.cu && ./mn_major_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe(const __half* __restrict__ a_img, const __half* __restrict__ b_img, float* __restrict__ d,
                                                uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* As = reinterpret_cast<__half*>(smem);                       // 16 col-groups x 2048 B = 32 KiB (128 cols)
    __half* Bs = reinterpret_cast<__half*>(smem + 32768);               // N/8 col-groups x 2048 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 65536);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x;
    for (int i = t; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = t; i < N * 256 / 16; i += 128) reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (t == 0) {
        // idesc: D=f32 (bit 4), A=B=f16, a_major = MN (bit 15), b_major = MN (bit 16), N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < 8; ++ks) {                                  // K = 16 rows per MMA
            const uint64_t ad = make_desc(smem_u32(As) + ks * 256, lbo, sbo);
            const uint64_t bd = make_desc(smem_u32(Bs) + ks * 256, lbo, sbo);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = t >> 5, lane = t & 31;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) d[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
    constexpr int N = 64;
    std::vector<__half> a(128 * 128), b(128 * N);
    std::vector<float> af(128 * 128), bf(128 * N);
    srand(1);
    auto img = [](int row, int col) { return (col / 8) * 1024 + row * 8 + (col % 8); };     // in halves
    for (int r = 0; r < 128; ++r) {
        for (int c = 0; c < 128; ++c) { float v = (rand() % 17 - 8) / 8.0f; af[r * 128 + c] = v; a[img(r, c)] = __float2half(v); }
        for (int c = 0; c < N; ++c) { float v = (rand() % 13 - 6) / 4.0f; bf[r * N + c] = v; b[img(r, c)] = __float2half(v); }
    }
    __half *ad, *bd; float* dd;
    cudaMalloc(&ad, a.size() * 2); cudaMalloc(&bd, b.size() * 2); cudaMalloc(&dd, 128 * N * 4);
    cudaMemcpy(ad, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(bd, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const uint32_t cfg[4][2] = {{128, 2048}, {2048, 128}, {256, 2048}, {2048, 256}};
    for (auto& c : cfg) {
        cudaMemset(dd, 0, 128 * N * 4);
        probe<N><<<1, 128, 100 * 1024>>>(ad, bd, dd, c[0], c[1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo %u sbo %u: CUDA error %s\n", c[0], c[1], cudaGetErrorString(e)); return 1; }
        std::vector<float> d(128 * N);
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0; int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int r = 0; r < 128; ++r) ref += (double)af[r * 128 + m] * bf[r * N + n];
                const double err = fabs(ref - d[m * N + n]);
                if (err > 1e-3) ++bad;
                if (err > worst) worst = err;
            }
        printf("MN-major A and B, LBO %4u SBO %4u: %s (worst abs err %.3g, %d of %d wrong)\n", c[0], c[1], bad ? "MISMATCH" : "OK", worst, bad, 128 * N);
    }
    return 0;
}

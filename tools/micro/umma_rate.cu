// Microbenchmark (B200, sm_100a): issue rate of tcgen05.mma kind::i8 (u8 x u8 -> s32, 128 x 256 x 32 per instruction) from
// static shared-memory operands — the ceiling of bq_umma.cu's filter pass, separated from its producers and epilogue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t a)
{
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void wait_bar(uint64_t *bar, unsigned parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// KIND 0: kind::i8 (u8), KIND 1: kind::f8f6f4 (e4m3) — same operand bytes, same shape
template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int kper, long long *cycles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (128 + N) * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 1);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    constexpr uint32_t IDESC_I8 = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC_F8 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // c = f32, a = b = e4m3 (0)
    if (tid == 0) {
        const uint32_t a = smem_u32(smem), b = a + 128 * 128;
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            for (int kk = 0; kk < kper; kk++) {
                const uint64_t da = desc(a + 32 * (kk & 3)), db = desc(b + 32 * (kk & 3));
                const uint32_t dcol = tm + (uint32_t)((it & 1) * N % 512);
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol), "l"(da),
                                 "l"(db), "r"(IDESC_I8), "r"(kk)
                                 : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol), "l"(da),
                                 "l"(db), "r"(IDESC_F8), "r"(kk)
                                 : "memory");
            }
            // a commit every `kper` MMAs on alternating barriers, waited for one iteration later (two accumulators in flight)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[it & 1])) : "memory");
            if (it >= 1) wait_bar(&bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
        }
        wait_bar(&bar[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int KIND, int N>
void run(const char *name)
{
    const int iters = 2000, kper = 48, grid = 148;
    long long *cyc;
    cudaMalloc(&cyc, grid * sizeof(long long));
    const size_t smem = (128 + N) * 128 + 1024;
    cudaFuncSetAttribute(k<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, N><<<grid, 128, smem>>>(10, kper, cyc);
    cudaEventRecord(e0);
    k<KIND, N><<<grid, 128, smem>>>(iters, kper, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c0 = 0;
    cudaMemcpy(&c0, cyc, sizeof c0, cudaMemcpyDeviceToHost);
    const double mmas = (double)iters * kper * grid, ops = mmas * 2.0 * 128 * N * 32;
    printf("%-40s %s  %.3f ms  %.1f cycles/MMA (SM 0)  %.0f TOP/s\n", name, cudaGetErrorString(e), ms, (double)c0 / ((double)iters * kper), ops / (ms * 1e-3) / 1e12);
    cudaFree(cyc);
}

int main()
{
    run<0, 256>("tcgen05.mma kind::i8     128x256x32");
    run<0, 128>("tcgen05.mma kind::i8     128x128x32");
    run<1, 256>("tcgen05.mma kind::f8f6f4 128x256x32");
    run<1, 128>("tcgen05.mma kind::f8f6f4 128x128x32");
    return 0;
}

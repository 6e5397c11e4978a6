// Microbenchmark (B200, sm_100a): issue rate of the legacy IMMA.16832.S8 path and of the bit->byte expansion that feeds it.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_rate imma_rate.cu && ./imma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0: IMMAs only (16 independent accumulators per warp). MODE 1: + the expansion ALU work of a 32x64 warp tile per k-step
// (12 words -> 24 registers, SHF + LOP3 each). MODE 2: same with IMAD-shift (fma pipe) + LOP3 (alu pipe).
template <int MODE>
__global__ void __launch_bounds__(256) k(const uint32_t *in, int *out, int iters)
{
    int c[16][4] = {};
    uint32_t w[12];
    for (int i = 0; i < 12; i++) w[i] = in[(threadIdx.x + 37 * i) & 1023];
    const int t = threadIdx.x & 3;
    for (int it = 0; it < iters; it++) {
        uint32_t a[2][4], b[8][2];
        if (MODE == 0) {
            for (int i = 0; i < 2; i++) for (int j = 0; j < 4; j++) a[i][j] = w[i * 4 + j];
            for (int i = 0; i < 8; i++) { b[i][0] = w[i]; b[i][1] = w[i + 4]; }
        } else if (MODE == 1) {
            for (int i = 0; i < 2; i++) {
                a[i][0] = (w[2 * i] >> (2 * t)) & 0x01010101u; a[i][2] = (w[2 * i] >> (2 * t + 1)) & 0x01010101u;
                a[i][1] = (w[2 * i + 1] >> (2 * t)) & 0x01010101u; a[i][3] = (w[2 * i + 1] >> (2 * t + 1)) & 0x01010101u;
            }
            for (int i = 0; i < 8; i++) { b[i][0] = (w[4 + i] >> (2 * t)) & 0x01010101u; b[i][1] = (w[4 + i] >> (2 * t + 1)) & 0x01010101u; }
        } else {
            const uint32_t m0 = 1u << (7 - 2 * t), m1 = 1u << (6 - 2 * t);
            for (int i = 0; i < 2; i++) {
                a[i][0] = (w[2 * i] * m0) & 0x80808080u; a[i][2] = (w[2 * i] * m1) & 0x80808080u;
                a[i][1] = (w[2 * i + 1] * m0) & 0x80808080u; a[i][3] = (w[2 * i + 1] * m1) & 0x80808080u;
            }
            for (int i = 0; i < 8; i++) { b[i][0] = (w[4 + i] * m0) & 0x80808080u; b[i][1] = (w[4 + i] * m1) & 0x80808080u; }
        }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) imma(c[i * 8 + j], a[i], b[j]);
        for (int i = 0; i < 12; i++) w[i] = w[i] * 1664525u + 1013904223u + (uint32_t)it;  // new operands every k-step
    }
    int s = 0;
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, const uint32_t *in, int *out, int ctas_per_sm)
{
    const int iters = 20000, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(in, out, 100);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double immas = (double)grid * 8 * 16 * iters;
    printf("%-28s ctas/SM=%d  %.3f ms  %.1f G IMMA/s  %.1f Tops (2*16*8*32 per IMMA)  -> 1536-bit pairs/s = %.1f G\n", name, ctas_per_sm, ms,
           immas / ms / 1e6, immas * 8192 / ms / 1e9, immas * 4096 / 1536 / ms / 1e6);
}

int main()
{
    uint32_t *in; int *out;
    cudaMalloc(&in, 4096); cudaMemset(in, 0x5a, 4096);
    cudaMalloc(&out, 148 * 4 * 256 * 4);
    for (int c = 1; c <= 2; c++) {
        run<0>("imma only", in, out, c);
        run<1>("imma + SHF/LOP3 expansion", in, out, c);
        run<2>("imma + IMAD/LOP3 expansion", in, out, c);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

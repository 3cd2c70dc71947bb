// Clean throughput probe: 16 independent min/max chains per thread in inline PTX
// (no helper arithmetic), s32 vs f32, 2-input vs 3-input.  ops/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#define CHAIN_I(k) asm volatile("min.s32 %0, %0, %1; max.s32 %1, %1, %0;" : "+r"(a[k]), "+r"(b[k]));
#define CHAIN_F(k) asm volatile("min.f32 %0, %0, %1; max.f32 %1, %1, %0;" : "+f"(x[k]), "+f"(y[k]));
#define CHAIN_I3(k) a[k] = __vimin3_s32(a[k], b[k], c3 + it); b[k] = __vimax3_s32(b[k], a[k], c3 - it);
#define CHAIN_F3(k) asm volatile("min.f32 %0, %0, %1, %2; max.f32 %1, %1, %0, %2;" : "+f"(x[k]), "+f"(y[k]) : "f"(f3));
// packed 16-bit forms: two values per instruction (candidates for an exact pre-filter in the epilogue)
#define CHAIN_S16X2(k) a[k] = (int)__vmins2((unsigned)a[k], (unsigned)b[k]); b[k] = (int)__vmaxs2((unsigned)b[k], (unsigned)a[k]);
#define CHAIN_S16X2_3(k) a[k] = (int)__vimin3_s16x2((unsigned)a[k], (unsigned)b[k], (unsigned)(c3 + it)); b[k] = (int)__vimax3_s16x2((unsigned)b[k], (unsigned)a[k], (unsigned)(c3 - it));
#define CHAIN_H2(k) { __half2 p = *reinterpret_cast<__half2 *>(&a[k]), q = *reinterpret_cast<__half2 *>(&b[k]); p = __hmin2(p, q); q = __hmax2(q, p); a[k] = *reinterpret_cast<int *>(&p); b[k] = *reinterpret_cast<int *>(&q); }
#define CHAIN_BF2(k) { __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162 *>(&a[k]), q = *reinterpret_cast<__nv_bfloat162 *>(&b[k]); p = __hmin2(p, q); q = __hmax2(q, p); a[k] = *reinterpret_cast<int *>(&p); b[k] = *reinterpret_cast<int *>(&q); }
#define REP8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

template <int OP>
__global__ void __launch_bounds__(256) probe(int *out, int iters, int seed) {
    int a[8], b[8]; float x[8], y[8];
    const int c3 = seed * 17; const float f3 = seed * 0.37f;
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * (k + 3) + seed; b[k] = a[k] ^ 0x5555; x[k] = a[k] * 0.25f; y[k] = b[k] * 0.5f; }
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) { REP8(CHAIN_I) }
        if (OP == 1) { REP8(CHAIN_F) }
        if (OP == 2) { REP8(CHAIN_I3) }
        if (OP == 3) { REP8(CHAIN_F3) }
        if (OP == 6) { REP8(CHAIN_S16X2) }
        if (OP == 7) { REP8(CHAIN_S16X2_3) }
        if (OP == 8) { REP8(CHAIN_H2) }
        if (OP == 9) { REP8(CHAIN_BF2) }
        if (OP == 10) { CHAIN_I(0) CHAIN_H2(1) CHAIN_I(2) CHAIN_H2(3) CHAIN_I(4) CHAIN_H2(5) CHAIN_I(6) CHAIN_H2(7) }   // do s32 and f16x2 min/max share a pipe?
        if (OP == 4) { CHAIN_I(0) CHAIN_F(0) CHAIN_I(1) CHAIN_F(1) CHAIN_I(2) CHAIN_F(2) CHAIN_I(3) CHAIN_F(3) }   // 8 int + 8 float ops
        if (OP == 5) { asm volatile("lop3.b32 %0, %0, %1, 31, 0x36;" : "+r"(a[0]) : "r"(b[0])); CHAIN_I(1) CHAIN_I(2) CHAIN_I(3)
                       asm volatile("lop3.b32 %0, %0, %1, 31, 0x36;" : "+r"(a[4]) : "r"(b[4])); CHAIN_I(5) CHAIN_I(6) CHAIN_I(7) }
    }
    int r = 0;
    for (int k = 0; k < 8; ++k) r += a[k] + b[k] + __float_as_int(x[k]) + __float_as_int(y[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char *name) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, iters = 1 << 13;
    int *d; cudaMalloc(&d, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); probe<OP><<<blocks, 256>>>(d, iters, r + 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double total = (double)blocks * 256 * iters * 16;
    printf("%-10s %8.3f ms  %.1f ops/clk/SM (nominal clock %d MHz)\n", name, best,
           total / (best * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
    cudaFree(d);
}

int main() { run<0>("min.s32"); run<1>("min.f32"); run<2>("min3.s32"); run<3>("min3.f32"); run<4>("mix s32+f32");
    // instructions/clk/SM; every packed instruction handles 2 values
    run<6>("vmin.s16x2"); run<7>("vmin3.s16x2"); run<8>("hmin2.f16"); run<9>("hmin2.bf16"); run<10>("mix s32+f16x2");
    return 0; }

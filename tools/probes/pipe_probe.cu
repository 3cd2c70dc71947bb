// Throughput probes for instructions whose pipe decides the design of the
// A-operand builder / epilogue: F2FP (float2 -> half2 pack), HADD2.F32-style
// half->float, F2I, VIMNMX, FMNMX, LOP3, IMAD.  Prints thread-level ops/clk/SM.
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) probe(unsigned *out, int iters, float fa, int ia) {
    float x0 = threadIdx.x * 0.5f + fa, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    int i0 = threadIdx.x + ia, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) {          // F2FP pack
            __half2 h0 = __floats2half2_rn(x0, x1), h1 = __floats2half2_rn(x2, x3);
            unsigned u0 = *reinterpret_cast<unsigned *>(&h0), u1 = *reinterpret_cast<unsigned *>(&h1);
            acc ^= u0 + u1; x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f;
        } else if (OP == 1) {   // F2I
            acc ^= (unsigned)__float2int_rz(x0) + (unsigned)__float2int_rz(x1) + (unsigned)__float2int_rz(x2) + (unsigned)__float2int_rz(x3);
            x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f;
        } else if (OP == 2) {   // VIMNMX
            i0 = min(i0, i1 + it); i1 = max(i1, i2 - it); i2 = min(i2, i3 + it); i3 = max(i3, i0 - it);
        } else if (OP == 3) {   // half2 -> float2
            __half2 h = *reinterpret_cast<__half2 *>(&i0);
            float2 f = __half22float2(h);
            x0 += f.x; x1 += f.y; i0 += 0x00010001;
        } else if (OP == 4) {   // FMNMX
            x0 = fminf(x0, x1 + 1.f); x1 = fmaxf(x1, x2 - 1.f); x2 = fminf(x2, x3 + 1.f); x3 = fmaxf(x3, x0 - 1.f);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(x0 + x1 + x2 + x3) + i0 + i1 + i2 + i3;
}

template <int OP>
void run(const char *name, double ops_per_iter) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, iters = 1 << 14;
    unsigned *d; cudaMalloc(&d, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); probe<OP><<<blocks, 256>>>(d, iters, 1.f, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double total = (double)blocks * 256 * iters * ops_per_iter;
    printf("%-10s %8.3f ms  %.3e ops/s  = %.1f ops/clk/SM at %d MHz (loop includes helper ops)\n", name, best,
           total / (best * 1e-3), total / (best * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
    cudaFree(d);
}

int main() {
    run<0>("F2FP.pack", 2); run<1>("F2I", 4); run<2>("VIMNMX", 4); run<3>("H2->F2", 1); run<4>("FMNMX", 4);
    return 0;
}

// Do non-FP64 instructions issue "for free" in the shadow of FP64 instructions (2 pipe cycles
// each), or does every instruction cost issue time on top?  Mixed streams, 16 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int V> __global__ void k(double *out, const double *in, int n)
{
    double x[8]; int q[8]; float f[8];
    for (int i = 0; i < 8; i++) { x[i] = in[threadIdx.x + i]; q[i] = threadIdx.x * 7 + i; f[i] = (float)in[threadIdx.x + 8 + i]; }
    const double y0 = in[300];
#pragma unroll 1
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[i] = fma(x[i], y0, 1e-7);                                    // full-rate DFMA form
                if (V == 1) q[i] = (q[i] ^ (q[i] >> 3)) + it;                      // 2-3 ALU ops per DFMA
                if (V == 2) q[i] = __shfl_xor_sync(0xffffffffu, q[i], 1);          // 1 SHFL per DFMA
                if (V == 3) f[i] = fmaf(f[i], 0.999f, 1e-3f);                      // 1 FFMA per DFMA
                if (V == 4) { q[i] = (q[i] ^ (q[i] >> 3)) + it; f[i] = fmaf(f[i], 0.999f, 1e-3f); }
                if (V == 5) q[i] = (x[i] > 1.0) ? q[i] + 1 : q[i] - 1;             // DSETP + SEL
            }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += x[i] + q[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(const char *name)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in; cudaMalloc(&out, (size_t)nsm * 512 * 8); cudaMalloc(&in, 8192);
    double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 0.999 + 1e-6 * i; cudaMemcpy(in, h, 8192, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int n = 4000;
    k<V><<<nsm, 512>>>(out, in, n / 8);
    cudaEventRecord(a); k<V><<<nsm, 512>>>(out, in, n); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double wi = (double)n * 32 * 16;  // DFMA warp-instr per SM
    printf("%-46s %.3f DFMA/clk/SM -> %.2f cycles per DFMA slot per SMSP\n", name, wi / (ms * 1e-3) / 1.965e9,
           4.0 / (wi / (ms * 1e-3) / 1.965e9));
}
int main()
{
    run<0>("DFMA only");
    run<1>("DFMA + 3 ALU (LOP/SHF/IADD)");
    run<2>("DFMA + 1 SHFL");
    run<3>("DFMA + 1 FFMA");
    run<4>("DFMA + 3 ALU + 1 FFMA");
    run<5>("DFMA + DSETP + 2 ALU");
    return 0;
}

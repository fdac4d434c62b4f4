// FP64 issue-rate vs operand pattern: does a DFMA with three distinct register operands issue
// slower than one with constant / reused operands?  (register-file bank pressure)
#include <cstdio>
#include <cuda_runtime.h>
template <int V> __global__ void k(double *out, const double *in, int n)
{
    double x[8], y[8], z[8];
    for (int i = 0; i < 8; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[64 + threadIdx.x + i]; z[i] = in[128 + threadIdx.x + i]; }
    const double y0 = y[0], z0 = z[0];
#pragma unroll 1
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (V == 0) x[i] = fma(x[i], y[i], z[i]);            // 3 distinct register operands
                if (V == 1) x[i] = fma(x[i], y0, z[i]);              // one multiplicand shared by all
                if (V == 2) x[i] = fma(x[i], y0, z0);                // two shared
                if (V == 3) x[i] = fma(x[i], 0.999999, 1e-7);        // immediates / constant bank
                if (V == 4) x[i] = x[i] + z[i];                      // DADD 2 regs
                if (V == 5) x[i] = x[i] * y[i];                      // DMUL 2 regs
                if (V == 6) x[i] = fma(x[i], x[i], z[i]);            // repeated operand
                if (V == 7) x[i] = fma(y[i], z[i], x[i]);            // accumulate form, 3 distinct
                if (V == 8) x[i] = fma(y[i], z[(i + 1) & 7], x[i]);  // accumulate, mixed banks
            }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(const char *name)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in; cudaMalloc(&out, (size_t)nsm * 512 * 8); cudaMalloc(&in, 4096);
    double h[512]; for (int i = 0; i < 512; i++) h[i] = 0.999 + 1e-6 * i; cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int n = 4000;
    k<V><<<nsm, 512>>>(out, in, n / 8);
    cudaEventRecord(a); k<V><<<nsm, 512>>>(out, in, n); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double wi = (double)n * 32 * 16;  // warp-instr per SM
    printf("%-44s %.3f warp-instr/clk/SM (@1965MHz)\n", name, wi / (ms * 1e-3) / 1.965e9);
}
int main()
{
    run<0>("DFMA x=fma(x,y[i],z[i]) 3 distinct regs");
    run<1>("DFMA x=fma(x,y0,z[i]) shared multiplicand");
    run<2>("DFMA x=fma(x,y0,z0) two shared");
    run<3>("DFMA x=fma(x,imm,imm)");
    run<4>("DADD x=x+z[i]");
    run<5>("DMUL x=x*y[i]");
    run<6>("DFMA x=fma(x,x,z[i])");
    run<7>("DFMA x=fma(y[i],z[i],x) 3 distinct");
    run<8>("DFMA x=fma(y[i],z[i+1],x) 3 distinct");
    return 0;
}

// Mixed instruction stream in the solver kernel's proportions (per 16 FP64: ~4.4 SHFL, 1 MUFU,
// 1.6 FSEL, 1 IMAD, 1 LOP/ISETP), all independent (ILP 8), 16 warps/SM: does the FP64 pipe stay
// saturated, i.e. do the other pipes overlap with it?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_seed(double x) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
template <int V> __global__ void k(double *out, const double *in, int n)
{
    double x[8], y[8], z[8]; int q[8];
    for (int i = 0; i < 8; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[64 + threadIdx.x + i]; z[i] = in[128 + threadIdx.x + i]; q[i] = threadIdx.x + i; }
    const double y0 = in[300];
#pragma unroll 1
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {          // 16 FP64 per r: 8 full-rate DFMA + 4 three-reg DFMA + 4 DMUL/DADD
                x[i] = fma(x[i], y0, 1e-7);
                if (i & 1) y[i] = fma(y[i], z[i], x[i]); else z[i] = z[i] * y0 + 0.0;
            }
            if (V >= 1) {
#pragma unroll
                for (int i = 0; i < 2; i++) z[i] = __shfl_xor_sync(0xffffffffu, z[i], 1 + r);   // 4 SHFL
            }
            if (V >= 2) { y[0] = rcp_seed(y[0]); }                                               // 1 MUFU
            if (V >= 3) { x[1] = (q[1] & 1) ? x[1] : z[1]; q[2] = q[2] * 3 + it; q[3] = (q[3] ^ it) & 0xffff; }  // 2 FSEL + IMAD + LOP3
        }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += x[i] + y[i] + z[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(const char *name, int warps)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in; cudaMalloc(&out, (size_t)nsm * 1024 * 8); cudaMalloc(&in, 8192);
    double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 0.999 + 1e-6 * i; cudaMemcpy(in, h, 8192, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int n = 8000;
    k<V><<<nsm, warps * 32>>>(out, in, n / 8);
    cudaEventRecord(a); k<V><<<nsm, warps * 32>>>(out, in, n); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fp64 = (double)n * 32 * (warps / 4.0);   // FP64 warp-instr per SMSP
    printf("%-52s %2d warps/SM: %.2f cycles per FP64 instr per SMSP\n", name, warps, (ms * 1e-3) * 1.965e9 / fp64);
}
int main()
{
    for (int w : {16, 4}) {
        run<0>("FP64 mix only (8 DFMA-fast, 4 DFMA-3reg, 4 DMUL)", w);
        run<1>("+ 4 SHFL per 16 FP64", w);
        run<2>("+ 4 SHFL + 1 MUFU", w);
        run<3>("+ 4 SHFL + 1 MUFU + 2 FSEL + IMAD + LOP3", w);
    }
    return 0;
}

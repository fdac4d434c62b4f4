// Micro-benchmarks that size the solver kernel's design choices on the actual B200:
// dependent-issue latencies (DFMA/DADD/DMUL, MUFU.RCP64H, SHFL), pipe throughputs, and the
// accuracy of Newton-refined reciprocals seeded by MUFU.RCP64H.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_seed(double x) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

template <int OP> __global__ void lat_kernel(double *out, long long *cyc, int n, double a, double b)
{
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if (OP == 0) x = fma(x, b, a);
            if (OP == 1) x = x + a;
            if (OP == 2) x = x * b;
            if (OP == 3) x = rcp_seed(x);
            if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1);
            if (OP == 5) { int y = __double2loint(x); y = __shfl_xor_sync(0xffffffffu, y, 1); x = __hiloint2double(__double2hiint(x), y); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> __global__ void tput_kernel(double *out, int n, double a, double b)
{
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = a + threadIdx.x * 1e-9 + k;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (OP == 0) x[k] = fma(x[k], b, a);
                if (OP == 3) x[k] = rcp_seed(x[k]);
                if (OP == 4) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1);
            }
    }
    double s = 0; for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void rcp_acc_kernel(double lo, double hi, int n, double *maxerr /*[5]*/)
{
    double e0 = 0, e3 = 0, e4 = 0, e5 = 0, e2 = 0;
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < n; i += nt) {
        double f = (i + 0.5) / n;
        double x = lo * pow(hi / lo, f);
        if (i & 1) x = -x;
        double ex = 1.0 / x;
        double r = rcp_seed(x);
        e0 = fmax(e0, fabs(r - ex) / fabs(ex));
        // 2 fma: one Newton step
        { double e = fma(-x, r, 1.0); double r2 = fma(r, e, r); e2 = fmax(e2, fabs(r2 - ex) / fabs(ex)); }
        // 3 fma: cubic
        double e = fma(-x, r, 1.0); double ec = fma(e, e, e); double r3 = fma(r, ec, r);
        e3 = fmax(e3, fabs(r3 - ex) / fabs(ex));
        // 4 fma: two Newton steps
        { double ea = fma(-x, r, 1.0); double ra = fma(r, ea, r); ea = fma(-x, ra, 1.0); ra = fma(ra, ea, ra);
          e4 = fmax(e4, fabs(ra - ex) / fabs(ex)); }
        // 5 fma: cubic + Newton
        double eb = fma(-x, r3, 1.0); double r5 = fma(r3, eb, r3);
        e5 = fmax(e5, fabs(r5 - ex) / fabs(ex));
    }
    // crude max reduce via atomics on ordered ints (positive doubles)
    atomicMax((unsigned long long *)&maxerr[0], (unsigned long long)__double_as_longlong(e0));
    atomicMax((unsigned long long *)&maxerr[1], (unsigned long long)__double_as_longlong(e2));
    atomicMax((unsigned long long *)&maxerr[2], (unsigned long long)__double_as_longlong(e3));
    atomicMax((unsigned long long *)&maxerr[3], (unsigned long long)__double_as_longlong(e4));
    atomicMax((unsigned long long *)&maxerr[4], (unsigned long long)__double_as_longlong(e5));
}

template <int OP> void run_lat(const char *name)
{
    double *out; long long *cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    lat_kernel<OP><<<1, 32>>>(out, cyc, 64, 1.000001, 0.999999);
    lat_kernel<OP><<<1, 32>>>(out, cyc, 256, 1.000001, 0.999999);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("latency  %-22s %.2f cycles/op\n", name, (double)c / (256.0 * 32));
    cudaFree(out); cudaFree(cyc);
}
template <int OP> void run_tput(const char *name, int warps_per_sm)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, (size_t)nsm * 1024 * 8 * 2);
    int n = 4000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    tput_kernel<OP><<<nsm, warps_per_sm * 32>>>(out, n / 4, 1.000001, 0.999999);
    cudaEventRecord(a);
    tput_kernel<OP><<<nsm, warps_per_sm * 32>>>(out, n, 1.000001, 0.999999);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)n * 32 * warps_per_sm;     // warp-instructions per SM
    printf("tput     %-22s %2d warps/SM: %.3f warp-instr/ns/SM (%.2f per clk at %d MHz nominal)\n", name, warps_per_sm,
           ops / (ms * 1e6), ops / (ms * 1e6) / (clk * 1e-6), clk / 1000);
    cudaFree(out);
}

int main()
{
    run_lat<0>("DFMA"); run_lat<1>("DADD"); run_lat<2>("DMUL"); run_lat<3>("MUFU.RCP64H");
    run_lat<4>("SHFL f64 (2x SHFL)"); run_lat<5>("SHFL b32");
    for (int w : {4, 8, 16, 32}) { run_tput<0>("DFMA", w); }
    for (int w : {4, 16}) { run_tput<3>("MUFU.RCP64H", w); run_tput<4>("SHFL f64 (2x SHFL)", w); }
    double *me; cudaMalloc(&me, 5 * 8);
    const double ranges[][2] = {{1e-3, 1e3}, {1e-30, 1e-20}, {1e20, 1e30}, {0.5, 2.0}};
    for (auto &r : ranges) {
        cudaMemset(me, 0, 40);
        rcp_acc_kernel<<<296, 256>>>(r[0], r[1], 1 << 24, me);
        double h[5]; cudaMemcpy(h, me, 40, cudaMemcpyDeviceToHost);
        printf("rcp accuracy x in [%g,%g]: seed %.3e | 2fma %.3e | 3fma(cubic) %.3e | 4fma %.3e | 5fma %.3e  (eps=%.3e)\n",
               r[0], r[1], h[0], h[1], h[2], h[3], h[4], 2.22e-16);
    }
    return 0;
}

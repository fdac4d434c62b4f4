// trpl_kernels.cu -- sm_100a kernels + C ABI of libtrpl_b200.so (see include/trpl_b200.h).
//
// Hot path replaced: pvSimPCR.py:14-401 (tEvol/iterate/pcreduce/norm2 + pvSim host driver) and
// probs.py:20-85 (kernel_lnP, log_kernel), plus the glue bayeslib.simulate runs between them
// (bayeslib.py:150-196).  Written from scratch for B200; nothing here is a translation of the
// numba kernels:
//
//   * one WARP owns one (sample, curve) simulation for its whole time integration; lane p holds
//     M consecutive grid nodes (L = 128 -> M = 4) of N, P, E and the BDF history sums in
//     REGISTERS; the four older BDF levels live in a lane-private shared-memory ring (no
//     barriers anywhere -- the reference round-trips the state through global memory every
//     step and needs ~50 __syncthreads per Newton iteration);
//   * each tridiagonal system is solved by a register-local partition sweep (M-1 interior rows
//     per lane eliminated in place) + a 32-lane parallel cyclic reduction on the interface
//     unknowns done with warp shuffles (5 normalised stages instead of the reference's
//     log2(L)-1 shared-memory stages over all L rows);
//   * the L1 residual norms of both species are reduced together with a 4-value butterfly;
//   * PL(t) is a warp reduction; 32 consecutive PL values are staged one per lane and then
//     consumed together: written with one coalesced store (trpl_solve_pl) and/or turned into
//     log10, time-interpolated onto the observation times and accumulated into the squared
//     log-residual sum (trpl_solve_loglik) -- PL never goes to HBM on the fused path;
//   * warps fetch work items from a global atomic counter (persistent CTAs), so a slowly
//     converging sample never idles a wave.
//
// All arithmetic is FP64 like the reference (pvSimPCR.py:11,113-125).  Divisions are replaced by
// a Newton-refined reciprocal (MUFU.RCP64H seed), accurate to <= 1 ulp; see DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>

#include "trpl_b200.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_CTA = 4;
#ifndef TRPL_MIN_CTAS
#define TRPL_MIN_CTAS 4
#endif

struct ObsDev {
    int n;
    int pad_;
    const int *hi;
    const double *whi;
    const double *wlo;
    const double *val;
};

struct CurveDev {
    double scales[TRPL_NPAR];
    double init_mul;      // dx^3, or 1 when the profile is already in grid units
    double redim;         // dx^2 * dt                         (pvSimPCR.py:393)
    const double *init;   // [L]
    void *pl_out;         // row base for sample 0, or nullptr
    int t_last;           // last time index to integrate to (<= T)
    int pad_;
    ObsDev obs[TRPL_MAX_EXP];
};

struct KArgs {
    const double *x;
    long long ldx;
    long long S;
    long long pl_stride;
    double TOL;
    double *sse;                 // [E][C][S] or nullptr
    int *status;                 // [C][S] or nullptr
    long long *iters;            // [C][S] or nullptr
    unsigned long long *counter; // work-item counter (zeroed by the host)
    int mag_col;                 // < 0: no magnitude offset
    int C, E, L, plT, max_iter, max_order, flags, pl_dtype;
    CurveDev curves[TRPL_MAX_CURVES];
};

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double rcp64(double x)
{
    // MUFU.RCP64H seed (rel. error <= 1e-6, measured on B200 with tools/microbench.cu) followed by
    // one cubically convergent step r*(1 + e + e^2): max error 1 ulp (2.2e-16, measured over 2^24
    // operands in four magnitude ranges).  No slow path: operands here are normal, finite and far
    // from the exponent limits.
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }

// ---------------------------------------------------------------------------------------------
// Communication among the lanes that share one simulation.  W = warps per simulation.
//   W == 1: warp shuffles / votes only (the production path for L <= 256).
//   W  > 1: one CTA of W warps per simulation (fine grids, L up to 128*W); values travel through a
//           ping-pong exchange buffer in shared memory, one __syncthreads per exchange.  Every
//           thread of the CTA executes the same sequence of exchanges.
// g = index of this lane among the 32*W lanes of the simulation.
// ---------------------------------------------------------------------------------------------
template <int W>
struct Comm {
    int g;            // lane index within the simulation
    double *xb;       // W > 1: exchange buffer [2][3][G] doubles
    double *red;      // W > 1: reduction scratch [2][W][4] doubles
    int phase;        // ping-pong selector of xb
    int rphase;       // ping-pong selector of red

    // K values from lane g-dm (-> vm) and lane g+dp (-> vp); out-of-range sources return the
    // caller's own value (always multiplied by an exact zero downstream).  For W > 1 the barrier
    // of the exchange also OR-reduces `busy` over the simulation (BAR.RED.OR) and returns it, so a
    // block-wide vote costs no barrier of its own; W == 1 returns true.
    template <int K, bool WANT_M, bool WANT_P>
    __device__ __forceinline__ bool xchg(const double (&v)[K], const int dm, const int dp,
                                         double (&vm)[K], double (&vp)[K], const bool busy = true)
    {
        if constexpr (W == 1) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (WANT_M) vm[k] = __shfl_up_sync(FULL, v[k], dm);
                if (WANT_P) vp[k] = __shfl_down_sync(FULL, v[k], dp);
            }
            return true;
        } else {
            constexpr int G = 32 * W;
            double *buf = xb + phase * (3 * G);
#pragma unroll
            for (int k = 0; k < K; k++) buf[k * G + g] = v[k];
            const bool any_busy = __syncthreads_or(busy) != 0;
            const int im = (g - dm >= 0) ? g - dm : g;
            const int ip = (g + dp < G) ? g + dp : g;
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (WANT_M) vm[k] = buf[k * G + im];
                if (WANT_P) vp[k] = buf[k * G + ip];
            }
            phase ^= 1;
            return any_busy;
        }
    }
    __device__ __forceinline__ double from_prev(const double v)     // value of lane g-1
    {
        double a[1] = {v}, m[1], p_[1];
        xchg<1, true, false>(a, 1, 1, m, p_);
        return m[0];
    }
    __device__ __forceinline__ double from_next(const double v)     // value of lane g+1
    {
        double a[1] = {v}, m[1], p_[1];
        xchg<1, false, true>(a, 1, 1, m, p_);
        return p_[0];
    }
    __device__ __forceinline__ bool all(const bool pred)
    {
        if constexpr (W == 1) return __all_sync(FULL, pred);
        else return __syncthreads_and(pred) != 0;
    }
    __device__ __forceinline__ double sum(double v)                 // total over the simulation
    {
        v = warp_sum(v);
        if constexpr (W > 1) {
            double *r = red + rphase * (W * 4);
            if ((threadIdx.x & 31) == 0) r[(threadIdx.x >> 5) * 4] = v;
            __syncthreads();
            v = 0.0;
#pragma unroll
            for (int w = 0; w < W; w++) v += r[w * 4];
            rphase ^= 1;
        }
        return v;
    }
    // stop rule errN < TOL and errP < TOL (pvSimPCR.py:213-216) with err = sum|res| / sum|b|, evaluated
    // division-free as z = sum(|res| - TOL*|b|) < 0 for both species (one 2-value butterfly).
    __device__ __forceinline__ void stop_rule(const double zN, const double zP, bool &converged,
                                              bool &nonfinite)
    {
        const int lane = threadIdx.x & 31;
        const bool hi16 = (lane & 16) != 0;
        double k = hi16 ? zP : zN;
        const double sd = hi16 ? zN : zP;
        k += __shfl_xor_sync(FULL, sd, 16);
        k += __shfl_xor_sync(FULL, k, 8);
        k += __shfl_xor_sync(FULL, k, 4);
        k += __shfl_xor_sync(FULL, k, 2);
        k += __shfl_xor_sync(FULL, k, 1);
        // lanes 0-15: zN of this warp, lanes 16-31: zP
        if constexpr (W > 1) {
            double *r = red + rphase * (W * 4);
            if ((lane & 15) == 0) r[(threadIdx.x >> 5) * 4 + (lane >> 4)] = k;
            __syncthreads();
            k = 0.0;
#pragma unroll
            for (int w = 0; w < W; w++) k += r[w * 4 + (lane >> 4)];
            rphase ^= 1;
        }
        converged = __all_sync(FULL, k < 0.0);
        nonfinite = __any_sync(FULL, !(fabs(k) <= DBL_MAX));
    }
};

// Tridiagonal solve, M rows per lane (row n = M*g + j):  l[j] x[n-1] + d[j] x[n] + u[j] x[n+1] = b[j].
// Rows outside the physical system must be identity rows (l=u=0, d=1).  l of the first row
// and u of the last physical row must be 0.
// Returns the new value of the previous lane's last node (needed by the callers anyway).
template <int M, int W>
__device__ __forceinline__ double tridiag_solve(const double (&l)[M], const double (&d)[M],
                                                const double (&u)[M], const double (&b)[M],
                                                double (&x)[M], Comm<W> &cm)
{
    double Lr, Dr, Ur, Br;
    double c[M > 1 ? M - 1 : 1], y[M > 1 ? M - 1 : 1], v[M > 1 ? M - 1 : 1], w[M > 1 ? M - 1 : 1];
    if constexpr (M > 1) {
        // interior rows 0..M-2:  x_j = y_j - v_j * s_left - w_j * s_own
        // Pivot reciprocals from the leading principal minors m_{j+1} = d_j m_j - l_j u_{j-1} m_{j-1}
        // (1/pivot_j = m_j / m_{j+1}): the M-1 reciprocals are independent of each other, so
        // their MUFU+Newton chains overlap instead of forming one serial chain.
        double ip[M - 1];
        {
            double mm[M];                 // mm[j] = m_{j+1}
            mm[0] = d[0];
            if constexpr (M > 2) mm[1] = fma(d[1], d[0], -(l[1] * u[0]));
#pragma unroll
            for (int j = 2; j < M - 1; j++) mm[j] = fma(d[j], mm[j - 1], -((l[j] * u[j - 1]) * mm[j - 2]));
            ip[0] = rcp64(mm[0]);
#pragma unroll
            for (int j = 1; j < M - 1; j++) ip[j] = mm[j - 1] * rcp64(mm[j]);
        }
        c[0] = u[0] * ip[0];
        y[0] = b[0] * ip[0];
        v[0] = l[0] * ip[0];
#pragma unroll
        for (int j = 1; j < M - 1; j++) {
            c[j] = u[j] * ip[j];
            y[j] = fma(-l[j], y[j - 1], b[j]) * ip[j];
            v[j] = (-l[j] * v[j - 1]) * ip[j];
        }
        w[M - 2] = c[M - 2];
#pragma unroll
        for (int j = M - 3; j >= 0; j--) {
            y[j] = fma(-c[j], y[j + 1], y[j]);
            v[j] = fma(-c[j], v[j + 1], v[j]);
            w[j] = -c[j] * w[j + 1];
        }
        // interface row (local M-1) couples s_left, s_own and the next lane's first interior row
        double mine[3] = {y[0], v[0], w[0]}, nm[3], nx[3];
        cm.template xchg<3, false, true>(mine, 1, 1, nm, nx);
        const double y0n = nx[0], v0n = nx[1], w0n = nx[2];
        const double lr = l[M - 1], ur = u[M - 1];
        Lr = -lr * v[M - 2];
        Dr = fma(-ur, v0n, fma(-lr, w[M - 2], d[M - 1]));
        Ur = -ur * w0n;
        Br = fma(-ur, y0n, fma(-lr, y[M - 2], b[M - 1]));
    } else {
        Lr = l[0]; Dr = d[0]; Ur = u[0]; Br = b[0];
    }
    // parallel cyclic reduction over the 32*W interface unknowns, unit diagonal
    {
        double inv = rcp64(Dr);
        Lr *= inv; Ur *= inv; Br *= inv;
    }
#pragma unroll
    for (int rf = 1; rf < 32 * W; rf <<= 1) {
        // Off-diagonals shrink quadratically per stage; once every |L|,|U| of the simulation is
        // below 2^-70 the remaining stages cannot change D = 1 or B in the last bit: stop.
        const int hl = __double2hiint(Lr) & 0x7fffffff, hu = __double2hiint(Ur) & 0x7fffffff;
        const bool busy = (rf < 2) || (max(hl, hu) >= ((1023 - 70) << 20));
        double mine[3] = {Lr, Ur, Br}, vm[3], vp[3];
        if constexpr (W == 1) {
            if (rf >= 2 && !__any_sync(FULL, busy)) break;
            cm.template xchg<3, true, true>(mine, rf, rf, vm, vp);
        } else {
            if (!cm.template xchg<3, true, true>(mine, rf, rf, vm, vp, busy)) break;
        }
        const double Lm = vm[0], Um = vm[1], Bm = vm[2];
        const double Lp = vp[0], Up = vp[1], Bp = vp[2];
        const double D = fma(-Lp, Ur, fma(-Um, Lr, 1.0));
        const double B = fma(-Bp, Ur, fma(-Bm, Lr, Br));
        const double Ln = -Lm * Lr;
        const double Un = -Up * Ur;
        const double inv = rcp64(D);
        Br = B * inv;
        Lr = Ln * inv;
        Ur = Un * inv;
    }
    x[M - 1] = Br;
    const double sl = cm.from_prev(Br);
    if constexpr (M > 1) {
#pragma unroll
        for (int j = 0; j < M - 1; j++) x[j] = fma(-w[j], Br, fma(-v[j], sl, y[j]));
    }
    return sl;
}

// lane-private ring of the 4 older BDF levels: [slot 0..3][field N,P,E][M doubles per lane]
template <int M>
struct Ring {
    double *base;   // warp base + lane offset
    // element (slot, field, j): chunks of 2 doubles per lane keep 16-byte accesses conflict-free
    __device__ __forceinline__ void load(int slot, int field, double (&h)[M]) const
    {
        if constexpr (M == 1) {
            h[0] = base[(slot * 3 + field) * 32];
        } else {
#pragma unroll
            for (int q = 0; q < M / 2; q++) {
                const double2 t = *reinterpret_cast<const double2 *>(
                    base + ((slot * 3 + field) * (M / 2) + q) * 64);
                h[2 * q] = t.x;
                h[2 * q + 1] = t.y;
            }
        }
    }
    __device__ __forceinline__ void store(int slot, int field, const double (&h)[M]) const
    {
        if constexpr (M == 1) {
            base[(slot * 3 + field) * 32] = h[0];
        } else {
#pragma unroll
            for (int q = 0; q < M / 2; q++)
                *reinterpret_cast<double2 *>(base + ((slot * 3 + field) * (M / 2) + q) * 64) =
                    make_double2(h[2 * q], h[2 * q + 1]);
        }
    }
};

struct WarpScratch {      // per-warp shared scratch touched once every 32 PL samples
    double sse[TRPL_MAX_EXP];
    int pos[TRPL_MAX_EXP];
};

// ---------------------------------------------------------------------------------------------
// one (sample, curve) simulation, executed by W warps (W == 1: one warp; W > 1: one CTA)
// ---------------------------------------------------------------------------------------------
template <int M, bool PAD, int W>
__device__ __forceinline__ void run_sim(const KArgs &a, const int c, const long long s,
                                        double *ring_warp, WarpScratch *ws, const int lane,
                                        Comm<W> &cm)
{
    constexpr int G = 32 * W;
    const int g = cm.g;                         // lane index within the simulation
    const bool io_warp = (W == 1) || (g < 32);  // the warp that stages, stores and scores PL
    const CurveDev &cv = a.curves[c];
    const int L = a.L;
    const int flags = a.flags;
    const bool emu32 = (flags & TRPL_F_EMULATE_F32) != 0;

    // ---- parameters: non-dimensionalise (pvSimPCR.py:327-331), one column per lane, then broadcast
    double mpl = 0.0;
    if (lane < TRPL_NPAR) mpl = a.x[s * a.ldx + lane] * cv.scales[lane];
    const double N0 = __shfl_sync(FULL, mpl, 0), P0 = __shfl_sync(FULL, mpl, 1);
    const double DN = __shfl_sync(FULL, mpl, 2), DP = __shfl_sync(FULL, mpl, 3);
    const double rate = __shfl_sync(FULL, mpl, 4);
    const double sr0 = __shfl_sync(FULL, mpl, 5), srL = __shfl_sync(FULL, mpl, 6);
    const double CN = __shfl_sync(FULL, mpl, 7), CP = __shfl_sync(FULL, mpl, 8);
    const double tauN = __shfl_sync(FULL, mpl, 9), tauP = __shfl_sync(FULL, mpl, 10);
    const double Lam = __shfl_sync(FULL, mpl, 11);
    const double N0P0 = N0 * P0;
    const double hDN = 0.5 * DN, hDP = 0.5 * DP;
    const double CN_N0P0 = CN * N0P0, CP_N0P0 = CP * N0P0;
    const double LamDP = Lam * DP, LamDN = Lam * DN, hLamDP = 0.5 * LamDP, hLamDN = 0.5 * LamDN;
    const double TOL = a.TOL;
    const double mag = (a.mag_col >= 0) ? a.x[s * a.ldx + a.mag_col] : 0.0;

    // ---- geometry of this lane
    const int last_lane = (L - 1) / M;         // lane owning node L-1 (at j = M-1 since L % M == 0)
    bool ev[M + 1];                            // edge m = M*g + j is an interior edge (1..L-1)
#pragma unroll
    for (int j = 0; j <= M; j++) {
        const int m = M * g + j;
        // exact-fit grids (L == M*G): only edge 0 (first lane) and edge L (last lane) are
        // boundaries, so the selects on the inner edges fold away at compile time
        ev[j] = PAD ? ((m >= 1) && (m <= L - 1)) : (j == 0 ? (g != 0) : (j == M ? (g != G - 1) : true));
    }
    bool nv[M];                                // node n = M*g + j exists
#pragma unroll
    for (int j = 0; j < M; j++) nv[j] = PAD ? (M * g + j < L) : true;
    // surface rows: lane 0 applies the front surface to j=0, last_lane the back surface to j=M-1
    const bool is_first = (g == 0), is_last = (g == last_lane);
    const double srf = is_first ? sr0 : (is_last ? srL : 0.0);

    // ---- initial state (pvSimPCR.py:339-362): N = N0 + dN, P = P0 + dN, E = 0
    double N[M], P[M], E[M];
#pragma unroll
    for (int j = 0; j < M; j++) {
        const int n = M * g + j;
        double dn = 0.0;
        if (n < L) dn = cv.init[n] * cv.init_mul;
        N[j] = nv[j] ? N0 + dn : 0.0;
        P[j] = nv[j] ? P0 + dn : 0.0;
        E[j] = 0.0;
    }
    Ring<M> ring;
    ring.base = ring_warp + ((M == 1) ? lane : 2 * lane);
    {
        double z[M];
#pragma unroll
        for (int j = 0; j < M; j++) z[j] = 0.0;
#pragma unroll
        for (int sl = 0; sl < 4; sl++)
#pragma unroll
            for (int f = 0; f < 3; f++) ring.store(sl, f, z);
    }
    if (io_warp && lane < TRPL_MAX_EXP) {
        ws->sse[lane] = 0.0;
        ws->pos[lane] = 0;
    }
    __syncwarp();

    // neighbour values carried across iterations and steps
    double Nl, Nr, Pl, Pr;
    {
        double mine[2] = {N[M - 1], P[M - 1]}, vm[2], vp[2];
        cm.template xchg<2, true, false>(mine, 1, 1, vm, vp);
        Nl = vm[0]; Pl = vm[1];
        double mine2[2] = {N[0], P[0]};
        cm.template xchg<2, false, true>(mine2, 1, 1, vm, vp);
        Nr = vp[0]; Pr = vp[1];
    }
    double En = 0.0;   // E on edge M*g + M (owned by the next lane)

    const double mLN0P0 = -(double)L * N0P0;   // pvSimPCR.py:278
    const int t_last = cv.t_last;
    const int plT = a.plT;
    const int n_pl = t_last / plT + 1;
    double keep = 0.0;          // PL sample staged in this lane
    double lp_carry = 0.0;      // log PL of the sample preceding the current block of 32
    double pl0 = 1.0;           // PL(t=0) for self-normalisation
    long long iters_total = 0;
    int status = 0;
    int pl_idx = 0;             // index of the next PL sample
    int t_next_pl = 0;

    // consume a block of `cnt` staged PL samples starting at index idx0
    auto flush = [&](const int idx0, const int cnt) {
        double val;
        if (emu32) {
            float f = (float)keep;             // value rounded on store into the f32 buffer
            f = f / (float)cv.redim;           // plI_main /= dx**2*dt in float32
            val = (double)f;
        } else {
            val = keep / cv.redim;
        }
        if (cv.pl_out != nullptr && lane < cnt) {
            if (a.pl_dtype == TRPL_F32)
                reinterpret_cast<float *>(cv.pl_out)[s * a.pl_stride + idx0 + lane] = (float)val;
            else
                reinterpret_cast<double *>(cv.pl_out)[s * a.pl_stride + idx0 + lane] = val;
        }
        if (a.E == 0) return;
        if (flags & TRPL_F_SELF_NORMALIZE) {
            if (idx0 == 0) pl0 = __shfl_sync(FULL, val, 0);
            val = emu32 ? (double)((float)val / (float)pl0) : val / pl0;
        }
        double lp = val;
        if (flags & TRPL_F_LOG_PL) {
            if (emu32) {
                float f = (float)val;
                if ((double)f < DBL_MIN) f = 0.0f;   // (float)sys.float_info.min == 0  (probs.py:72-73)
                lp = (double)log10f(f);
            } else {
                lp = log10(val < DBL_MIN ? DBL_MIN : val);
            }
        }
        for (int e = 0; e < a.E; e++) {
            const ObsDev &ob = cv.obs[e];
            int pos = ws->pos[e];
            double acc = 0.0;
            for (;;) {
                const int i = pos + lane;
                const int h = (i < ob.n) ? ob.hi[i] : INT_MAX;
                const bool mine = h < idx0 + cnt;
                const unsigned bm = __ballot_sync(FULL, mine);
                if (bm == 0u) break;
                const int shi = mine ? h - idx0 : 0;       // 0..cnt-1
                const int slo = shi - 1;                   // -1..cnt-2
                const double y_hi = __shfl_sync(FULL, lp, shi & 31);
                double y_lo = __shfl_sync(FULL, lp, slo & 31);
                if (slo < 0) y_lo = lp_carry;
                double sq = 0.0;
                if (mine) {
                    // scipy interp1d._call_linear: w_hi*y_hi + w_lo*y_lo, no contraction
                    const double yi = __dadd_rn(__dmul_rn(ob.whi[i], y_hi), __dmul_rn(ob.wlo[i], y_lo));
                    double err = yi + mag;                  // probs.py:33-38
                    err -= ob.val[i];
                    sq = err * err;
                }
                acc += warp_sum(sq);
                const int took = __popc(bm);
                pos += took;
                if (took < 32) break;
            }
            if (lane == 0) {
                ws->pos[e] = pos;
                ws->sse[e] += acc;
            }
        }
        lp_carry = __shfl_sync(FULL, lp, cnt - 1);
        __syncwarp();
    };

    // =========================================================================================
    // time loop (pvSimPCR.py:237-293): t = 0 .. t_last, PL(t) emitted from the state at time t
    // =========================================================================================
    bool failed = false;
    int t;
    for (t = 0; t <= t_last; t++) {
        // ---- PL(t) = rate * (sum_n N*P - L*N0*P0)                          (pvSimPCR.py:276-281)
        bool emitted = false;
        if (t == t_next_pl) {
            emitted = true;
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < M; j++) part = fma(N[j], P[j], part);
            const double tot = cm.sum(part);
            const double plraw = rate * (tot + mLN0P0);
            if (lane == (pl_idx & 31)) keep = plraw;
            t_next_pl += plT;
            pl_idx++;
        }

        // ---- BDF coefficients, order ramp 1..5                            (pvSimPCR.py:241-250)
        double a0, a1, a2, a3, a4, a5;
        {
            int order = t + 1;
            if (order > 5) order = 5;
            if (order > a.max_order) order = a.max_order;
            a2 = a3 = a4 = a5 = 0.0;
            if (order == 1) { a0 = 1.0; a1 = -1.0; }
            else if (order == 2) { a0 = 1.5; a1 = -2.0; a2 = 0.5; }
            else if (order == 3) { a0 = 11.0 / 6; a1 = -3.0; a2 = 1.5; a3 = -1.0 / 3; }
            else if (order == 4) { a0 = 25.0 / 12; a1 = -4.0; a2 = 3.0; a3 = -4.0 / 3; a4 = 0.25; }
            else { a0 = 137.0 / 60; a1 = -5.0; a2 = 5.0; a3 = -10.0 / 3; a4 = 1.25; a5 = -0.2; }
        }

        // ---- history sums bU = a1 U(t) + a2 U(t-1) + ... + a5 U(t-4)       (pvSimPCR.py:133-135)
        double bN[M], bP[M], bE[M];
#pragma unroll
        for (int j = 0; j < M; j++) {
            bN[j] = a1 * N[j];
            bP[j] = a1 * P[j];
            bE[j] = a1 * E[j];
        }
        {
            const double ac[4] = {a2, a3, a4, a5};
#pragma unroll
            for (int i = 1; i <= 4; i++) {
                const int slot = (t - i) & 3;
                double h[M];
                ring.load(slot, 0, h);
#pragma unroll
                for (int j = 0; j < M; j++) bN[j] = fma(ac[i - 1], h[j], bN[j]);
                ring.load(slot, 1, h);
#pragma unroll
                for (int j = 0; j < M; j++) bP[j] = fma(ac[i - 1], h[j], bP[j]);
                ring.load(slot, 2, h);
#pragma unroll
                for (int j = 0; j < M; j++) bE[j] = fma(ac[i - 1], h[j], bE[j]);
            }
            const int slot = t & 3;    // level t replaces level t-4
            ring.store(slot, 0, N);
            ring.store(slot, 1, P);
            ring.store(slot, 2, E);
        }

        // ---- Newton / Gauss-Seidel iteration                              (pvSimPCR.py:147-216)
        int it = 0;
        bool nonfinite = false;
        for (;;) {
            double l[M], d[M], u[M], b[M];
            double zN = 0.0, zP = 0.0;    // sum(|residual| - TOL*|b|): err < TOL  <=>  z < 0
            bool converged_now = false, nonfinite_now = false;

            // ======== N system (P, E frozen) ========
            {
                double cu[M + 1], cl[M + 1];   // edge coefficients: cu[m] = upper of row m-1, cl[m] = lower of row m
#pragma unroll
                for (int j = 0; j <= M; j++) {
                    const double Ej = (j < M) ? E[j] : En;
                    cu[j] = sel(ev[j], fma(-hDN, Ej, -DN), 0.0);     // DN*(-E/2 - 1)
                    cl[j] = sel(ev[j], fma(hDN, Ej, -DN), 0.0);      // DN*(+E/2 - 1)
                }
#pragma unroll
                for (int j = 0; j < M; j++) {
                    const double Nj = N[j], Pj = P[j];
                    const double tp = fma(Nj, tauP, Pj * tauN);
                    const double NP = Nj * Pj;
                    const double npp = NP - N0P0;
                    const double r = rcp64(tp);
                    const double q = fma(-tauP, npp, Pj * tp);
                    const double srh = (q * r) * r;
                    const double cnN = CN * Nj;
                    const double aug = fma(Pj, fma(CP, Pj, cnN + cnN), -CN_N0P0);   // CN*N*P + CP*P^2 + CN*np
                    const double nds = fma(rate, Pj, srh) + aug;             // = -ds
                    l[j] = cl[j];
                    u[j] = cu[j + 1];
                    d[j] = ((a0 - cu[j]) - cl[j + 1]) + nds;
                    const double g = fma(CP, Pj, cnN) + (rate + r);
                    b[j] = fma(nds, Nj, -fma(g, npp, bN[j]));
                }
                // surface recombination rows                                 (pvSimPCR.py:164-170)
                {
                    const double Ns = is_first ? N[0] : N[M - 1];
                    const double Ps = is_first ? P[0] : P[M - 1];
                    const double rs = rcp64(Ns + Ps);
                    const double nd = (srf * fma(Ps, Ps, N0P0)) * (rs * rs);        // = -ds0
                    const double db = fma(-nd, Ns, (srf * fma(Ns, Ps, -N0P0)) * rs);
                    d[0] += is_first ? nd : 0.0;
                    b[0] -= is_first ? db : 0.0;
                    d[M - 1] += is_last ? nd : 0.0;
                    b[M - 1] -= is_last ? db : 0.0;
                }
                if (PAD) {
#pragma unroll
                    for (int j = 0; j < M; j++) {
                        l[j] = sel(nv[j], l[j], 0.0);
                        u[j] = sel(nv[j], u[j], 0.0);
                        d[j] = sel(nv[j], d[j], 1.0);
                        b[j] = sel(nv[j], b[j], 0.0);
                    }
                }
                // L1 residual of the current iterate                         (pvSimPCR.py:172, :14-40)
#pragma unroll
                for (int j = 0; j < M; j++) {
                    const double xm = (j == 0) ? Nl : N[j - 1];
                    const double xp = (j == M - 1) ? Nr : N[j + 1];
                    const double res = fma(l[j], xm, fma(d[j], N[j], fma(u[j], xp, -b[j])));
                    zN = fma(-TOL, fabs(b[j]), zN + fabs(res));
                }
                Nl = tridiag_solve<M, W>(l, d, u, b, N, cm);
                Nr = cm.from_next(N[0]);
            }

            // ======== P system (new N) ========
            {
                double cu[M + 1], cl[M + 1];
#pragma unroll
                for (int j = 0; j <= M; j++) {
                    const double Ej = (j < M) ? E[j] : En;
                    cu[j] = sel(ev[j], fma(hDP, Ej, -DP), 0.0);      // DP*(+E/2 - 1)
                    cl[j] = sel(ev[j], fma(-hDP, Ej, -DP), 0.0);     // DP*(-E/2 - 1)
                }
#pragma unroll
                for (int j = 0; j < M; j++) {
                    const double Nj = N[j], Pj = P[j];
                    const double tp = fma(Nj, tauP, Pj * tauN);
                    const double NP = Nj * Pj;
                    const double npp = NP - N0P0;
                    const double r = rcp64(tp);
                    const double q = fma(-tauN, npp, Nj * tp);
                    const double srh = (q * r) * r;
                    const double cpP = CP * Pj;
                    const double aug = fma(Nj, fma(CN, Nj, cpP + cpP), -CP_N0P0);   // CP*N*P + CN*N^2 + CP*np
                    const double nds = fma(rate, Nj, srh) + aug;
                    l[j] = cl[j];
                    u[j] = cu[j + 1];
                    d[j] = ((a0 - cu[j]) - cl[j + 1]) + nds;
                    const double g = fma(CN, Nj, cpP) + (rate + r);
                    b[j] = fma(nds, Pj, -fma(g, npp, bP[j]));
                }
                {
                    const double Ns = is_first ? N[0] : N[M - 1];
                    const double Ps = is_first ? P[0] : P[M - 1];
                    const double rs = rcp64(Ns + Ps);
                    const double nd = (srf * fma(Ns, Ns, N0P0)) * (rs * rs);
                    const double db = fma(-nd, Ps, (srf * fma(Ns, Ps, -N0P0)) * rs);
                    d[0] += is_first ? nd : 0.0;
                    b[0] -= is_first ? db : 0.0;
                    d[M - 1] += is_last ? nd : 0.0;
                    b[M - 1] -= is_last ? db : 0.0;
                }
                if (PAD) {
#pragma unroll
                    for (int j = 0; j < M; j++) {
                        l[j] = sel(nv[j], l[j], 0.0);
                        u[j] = sel(nv[j], u[j], 0.0);
                        d[j] = sel(nv[j], d[j], 1.0);
                        b[j] = sel(nv[j], b[j], 0.0);
                    }
                }
#pragma unroll
                for (int j = 0; j < M; j++) {
                    const double xm = (j == 0) ? Pl : P[j - 1];
                    const double xp = (j == M - 1) ? Pr : P[j + 1];
                    const double res = fma(l[j], xm, fma(d[j], P[j], fma(u[j], xp, -b[j])));
                    zP = fma(-TOL, fabs(b[j]), zP + fabs(res));
                }
                // ---- stop decision for THIS iteration (pvSimPCR.py:213-216): both L1 residuals are
                // known here, before the P solve; reducing them now lets the shuffle chain overlap
                // the solve.
                cm.stop_rule(zN, zP, converged_now, nonfinite_now);
                Pl = tridiag_solve<M, W>(l, d, u, b, P, cm);
                Pr = cm.from_next(P[0]);
            }

            // ======== E update on interior edges                           (pvSimPCR.py:205-209)
#pragma unroll
            for (int j = 0; j < M; j++) {
                const double Nm = (j == 0) ? Nl : N[j - 1];
                const double Pm = (j == 0) ? Pl : P[j - 1];
                const double den = fma(hLamDP, P[j] + Pm, fma(hLamDN, N[j] + Nm, a0));
                const double num = fma(LamDP, P[j] - Pm, fma(-LamDN, N[j] - Nm, -bE[j]));
                E[j] = sel(ev[j], num * rcp64(den), 0.0);
            }
            En = cm.from_next(E[0]);

            // ======== stop rule (pvSimPCR.py:213-216): decided by the flags computed before the P solve
            it++;
            if (nonfinite_now) { nonfinite = true; break; }
            if (converged_now) break;
            if (it >= a.max_iter) break;
        }
        iters_total += it;
        if (nonfinite || it >= a.max_iter) {                 // pvSimPCR.py:269-274
            status |= nonfinite ? TRPL_ST_NONFINITE : TRPL_ST_NOCONV;
            failed = true;
            // the reference stops before emitting PL(t): un-count the sample staged for this step
            if (emitted) pl_idx--;
            break;
        }
        if (io_warp && emitted && (pl_idx & 31) == 0) flush(pl_idx - 32, 32);
    }

    // ---- tail: partially filled block; after a failure everything from pl_idx on is NaN
    if (io_warp && (pl_idx & 31)) flush(pl_idx & ~31, pl_idx & 31);
    if (io_warp && failed && cv.pl_out != nullptr) {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        for (int i = pl_idx + lane; i < n_pl; i += 32) {
            if (a.pl_dtype == TRPL_F32)
                reinterpret_cast<float *>(cv.pl_out)[s * a.pl_stride + i] = (float)qnan;
            else
                reinterpret_cast<double *>(cv.pl_out)[s * a.pl_stride + i] = qnan;
        }
    }

    // ---- results
    __syncwarp();
    if (io_warp && lane == 0) {
        const long long cs = (long long)c * a.S + s;
        if (a.status) a.status[cs] = status;
        if (a.iters) a.iters[cs] = iters_total;
    }
    if (io_warp && a.sse != nullptr && lane < a.E) {
        double v = ws->sse[lane];
        if (failed && ws->pos[lane] < cv.obs[lane].n) v = __longlong_as_double(0x7ff8000000000000LL);
        a.sse[((long long)lane * a.C + c) * a.S + s] = v;
    }
    __syncwarp();
}

template <int M, bool PAD>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, TRPL_MIN_CTAS)
trpl_sim_kernel(const __grid_constant__ KArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ WarpScratch scratch[WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *ring_warp = smem + (size_t)warp * (4 * 3 * M * 32);
    const unsigned long long total = (unsigned long long)a.S * (unsigned long long)a.C;
    Comm<1> cm;
    cm.g = lane; cm.xb = nullptr; cm.red = nullptr; cm.phase = 0; cm.rphase = 0;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1ULL);
        item = __shfl_sync(FULL, item, 0);
        if (item >= total) break;
        const long long s = (long long)(item / (unsigned)a.C);
        const int c = (int)(item % (unsigned)a.C);
        run_sim<M, PAD, 1>(a, c, s, ring_warp, &scratch[warp], lane, cm);
    }
}

// Fine grids: one CTA of W warps per simulation, 4 nodes per lane (L <= 128*W).
template <int W>
__global__ void __launch_bounds__(W * 32, 16 / W)
trpl_sim_cta_kernel(const __grid_constant__ KArgs a)
{
    constexpr int M = 4;
    extern __shared__ __align__(16) double smem[];
    __shared__ WarpScratch scratch;
    __shared__ unsigned long long next_item;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *ring_warp = smem + (size_t)warp * (4 * 3 * M * 32);
    Comm<W> cm;
    cm.g = threadIdx.x;
    cm.xb = smem + (size_t)W * (4 * 3 * M * 32);
    cm.red = cm.xb + 2 * 3 * 32 * W;
    cm.phase = 0; cm.rphase = 0;
    const unsigned long long total = (unsigned long long)a.S * (unsigned long long)a.C;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next_item = atomicAdd(a.counter, 1ULL);
        __syncthreads();
        const unsigned long long item = next_item;
        if (item >= total) break;
        const long long s = (long long)(item / (unsigned)a.C);
        const int c = (int)(item % (unsigned)a.C);
        run_sim<M, true, W>(a, c, s, ring_warp, &scratch, lane, cm);
    }
}

// lnl[e][s] -= sum_c sse[e][c][s] (curves in order), status[s] = OR_c status[c][s]
__global__ void trpl_finish_kernel(const double *sse, double *lnl, const int *status_cs,
                                   int *status_s, long long S, int C, int E)
{
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int e = 0; e < E; e++) {
        double p = lnl[e * S + s];
        for (int c = 0; c < C; c++) p -= sse[((long long)e * C + c) * S + s];
        lnl[e * S + s] = p;
    }
    if (status_s) {
        int st = 0;
        for (int c = 0; c < C; c++) st |= status_cs[(long long)c * S + s];
        status_s[s] = st;
    }
}

// ---- probs.fastlog / log_kernel (probs.py:64-85) ---------------------------------------------
// HBM-bound streaming kernels: 4 independent loads in flight per thread (memory-level parallelism).
__global__ void trpl_log10_kernel_f64(double *x, long long n, double mn)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double v0 = x[i], v1 = x[i + stride], v2 = x[i + 2 * stride], v3 = x[i + 3 * stride];
        v0 = v0 < mn ? mn : v0; v1 = v1 < mn ? mn : v1; v2 = v2 < mn ? mn : v2; v3 = v3 < mn ? mn : v3;
        x[i] = log10(v0); x[i + stride] = log10(v1); x[i + 2 * stride] = log10(v2); x[i + 3 * stride] = log10(v3);
    }
    for (; i < n; i += stride) {
        double v = x[i];
        if (v < mn) v = mn;
        x[i] = log10(v);
    }
}
__device__ __forceinline__ float log10_clamp_f32(float v, double mn)
{
    if ((double)v < mn) v = (float)mn;
    return log10f(v);
}
__global__ void trpl_log10_kernel_f32(float *x, long long n, double mn)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const float v0 = x[i], v1 = x[i + stride], v2 = x[i + 2 * stride], v3 = x[i + 3 * stride];
        x[i] = log10_clamp_f32(v0, mn); x[i + stride] = log10_clamp_f32(v1, mn);
        x[i + 2 * stride] = log10_clamp_f32(v2, mn); x[i + 3 * stride] = log10_clamp_f32(v3, mn);
    }
    for (; i < n; i += stride) x[i] = log10_clamp_f32(x[i], mn);
}

// ---- probs.prob / kernel_lnP (probs.py:20-62): one warp per sample, coalesced row reads -------
__global__ void trpl_lnp_kernel(double *P, const double *pl, long long S, long long n, long long ld,
                                const double *values, const double *mag)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long j = warp; j < S; j += nwarps) {
        const double m = mag[j];
        const double *row = pl + j * ld;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        long long i = lane;
        for (; i + 96 < n; i += 128) {
            const double r0 = row[i], r1 = row[i + 32], r2 = row[i + 64], r3 = row[i + 96];
            const double e0 = (r0 + m) - values[i], e1 = (r1 + m) - values[i + 32];
            const double e2 = (r2 + m) - values[i + 64], e3 = (r3 + m) - values[i + 96];
            a0 = fma(e0, e0, a0); a1 = fma(e1, e1, a1); a2 = fma(e2, e2, a2); a3 = fma(e3, e3, a3);
        }
        for (; i < n; i += 32) {
            const double e = (row[i] + m) - values[i];
            a0 = fma(e, e, a0);
        }
        const double acc = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) P[j] += (0.0 - acc);
    }
}

// ---- shard-local log-sum-exp pieces (Visualization/utils.py:157-166) --------------------------
__global__ void trpl_lse_max_kernel(const double *x, long long n, double *out)
{
    __shared__ double sh[32];
    double m = -INFINITY;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = x[i];
        if (v == v && v > m) m = v;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
        if (threadIdx.x == 0) {
            // atomic max on doubles via ordered-integer trick
            unsigned long long *addr = reinterpret_cast<unsigned long long *>(out);
            unsigned long long old = *addr, assumed;
            do {
                assumed = old;
                if (__longlong_as_double((long long)assumed) >= m) break;
                old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(m));
            } while (assumed != old);
        }
    }
}
__global__ void trpl_lse_sum_kernel(const double *x, long long n, double *out)
{
    __shared__ double sh[32];
    const double mx = out[0];
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = x[i];
        if (v == v) acc += exp(v - mx);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) atomicAdd(out + 1, acc);
    }
}
__global__ void trpl_lse_init_kernel(double *out)
{
    out[0] = -INFINITY;
    out[1] = 0.0;
}

// ---- sample generation on the device (bayeslib.random_grid / make_grid, bayeslib.py:18-76) -----
// Counter-based Philox4x32-10: sample s, column j uses counter (s_lo, s_hi, j, 0) and key (seed_lo,
// seed_hi); u = 53 random bits / 2^53.  Same bounds / log / override semantics as the reference.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                              unsigned k0, unsigned k1, unsigned (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct GridArgs {
    double lo[16], hi[16];
    int do_log[16];
    int ncol, eq_mu, eq_s, eq_auger;
};

__global__ void trpl_random_grid_kernel(double *x, long long S, long long ldx, const GridArgs ga,
                                        unsigned long long seed, unsigned long long first)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long total = S * ga.ncol;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long s = i / ga.ncol;
        const int j = (int)(i - s * ga.ncol);
        int src = j;                                   // override_equal_*: copy the draw of another column
        if (ga.eq_mu && j == 2) src = 3;
        if (ga.eq_s && j == 6) src = 5;
        if (ga.eq_auger && j == 8) src = 7;
        const unsigned long long id = first + (unsigned long long)s;
        unsigned r[4];
        philox4x32_10((unsigned)id, (unsigned)(id >> 32), (unsigned)src, 0u, (unsigned)seed,
                      (unsigned)(seed >> 32), r);
        const unsigned long long bits = (((unsigned long long)r[0] << 32) | r[1]) >> 11;
        const double u = (double)bits * (1.0 / 9007199254740992.0);
        const double lo = ga.lo[src], hi = ga.hi[src];
        double v;
        if (lo == hi) v = lo;
        else if (ga.do_log[src]) {
            const double a = log10(lo), b = log10(hi);
            v = exp10(a + (b - a) * u);
        } else v = lo + (hi - lo) * u;
        x[s * ldx + j] = v;
    }
}

// ---- posterior products (Visualization/utils.py:157-285) -------------------------------------
__global__ void trpl_weights_kernel(const double *lnp, long long n, double lse, double *w)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = lnp[i];
        w[i] = (v == v) ? exp(v - lse) : 0.0;
    }
}

// numpy.histogram bin of v over nb uniform bins on [lo, hi] (right edge inclusive), or -1
__device__ __forceinline__ int hist_bin(double v, double lo, double hi, int nb)
{
    if (!(v >= lo) || !(v <= hi)) return -1;
    int b = (int)((v - lo) / (hi - lo) * nb);
    if (b >= nb) b = nb - 1;
    // guard the edges against rounding of the scaled position (numpy does the same correction)
    const double e0 = lo + (hi - lo) * b / nb, e1 = lo + (hi - lo) * (b + 1) / nb;
    if (v < e0 && b > 0) b--;
    else if (v >= e1 && b < nb - 1) b++;
    return b;
}

__global__ void trpl_hist_kernel(const double *x, long long ldx, int colx, int coly, const double *w,
                                 long long n, double lox, double hix, int nbx, double loy, double hiy,
                                 int nby, double *hist)
{
    extern __shared__ double sh[];
    const int nb = nbx * (coly >= 0 ? nby : 1);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const int bx = hist_bin(x[i * ldx + colx], lox, hix, nbx);
        int b = bx;
        if (coly >= 0) {
            const int by = hist_bin(x[i * ldx + coly], loy, hiy, nby);
            b = (bx < 0 || by < 0) ? -1 : bx * nby + by;
        }
        if (b >= 0) atomicAdd(&sh[b], w ? w[i] : 1.0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (sh[i] != 0.0) atomicAdd(&hist[i], sh[i]);
}

// out[0] = sum w, out[1+j] = sum w x_j, out[1+ncol+j*ncol+k] = sum w x_j x_k
// One thread per sample row (coalescing comes from the 32 rows of a warp being adjacent in
// memory and ncol <= 15 columns being read in order); accumulators live in shared memory per
// warp and are reduced with shuffles -> few atomics.
__global__ void trpl_moments_kernel(const double *x, long long ldx, int ncol, const double *w,
                                    long long n, double *out)
{
    const int nacc = 1 + ncol + ncol * ncol;
    extern __shared__ double acc_sh[];          // [nacc] per block
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc_sh[i] = 0.0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    double sw = 0.0, sx[15], sxx[15];           // per thread: sum w, sum w x_j, and ONE row of x x^T at a time
#pragma unroll
    for (int j = 0; j < 15; j++) { sx[j] = 0.0; sxx[j] = 0.0; }
    // pass structure: for each j0, accumulate sum w x_j0 x_k for all k (re-reading the row from L1/L2)
    for (int j0 = -1; j0 < ncol; j0++) {
#pragma unroll
        for (int k = 0; k < 15; k++) sxx[k] = 0.0;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
            const double ww = w[i];
            const double *row = x + i * ldx;
            if (j0 < 0) {
                sw += ww;
#pragma unroll
                for (int k = 0; k < 15; k++) if (k < ncol) sx[k] = fma(ww, row[k], sx[k]);
            } else {
                const double wx = ww * row[j0];
#pragma unroll
                for (int k = 0; k < 15; k++) if (k < ncol) sxx[k] = fma(wx, row[k], sxx[k]);
            }
        }
        if (j0 < 0) {
            sw = warp_sum(sw);
            if ((threadIdx.x & 31) == 0) atomicAdd(&acc_sh[0], sw);
#pragma unroll
            for (int k = 0; k < 15; k++) if (k < ncol) {
                const double t = warp_sum(sx[k]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&acc_sh[1 + k], t);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 15; k++) if (k < ncol) {
                const double t = warp_sum(sxx[k]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&acc_sh[1 + ncol + j0 * ncol + k], t);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += blockDim.x)
        if (acc_sh[i] != 0.0) atomicAdd(&out[i], acc_sh[i]);
}

// Single-pass variant for the usual sample matrix (NC columns known at compile time): one thread per
// row, sum w / sum w x_j / upper triangle of sum w x_j x_k in registers, one warp reduction at the end.
template <int NC>
__global__ void __launch_bounds__(128, 2)
trpl_moments_kernel_fixed(const double *x, long long ldx, const double *w, long long n, double *out)
{
    constexpr int NT = NC * (NC + 1) / 2;
    double sw = 0.0, sx[NC], sxx[NT];
#pragma unroll
    for (int j = 0; j < NC; j++) sx[j] = 0.0;
#pragma unroll
    for (int j = 0; j < NT; j++) sxx[j] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ww = w[i];
        double r[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) r[j] = x[i * ldx + j];
        sw += ww;
        int q = 0;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double wx = ww * r[j];
            sx[j] += wx;
#pragma unroll
            for (int k = j; k < NC; k++) sxx[q++] = fma(wx, r[k], sxx[q]);
        }
    }
    const bool lead = (threadIdx.x & 31) == 0;
    sw = warp_sum(sw);
    if (lead && sw != 0.0) atomicAdd(&out[0], sw);
    int q = 0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const double t = warp_sum(sx[j]);
        if (lead && t != 0.0) atomicAdd(&out[1 + j], t);
#pragma unroll
        for (int k = j; k < NC; k++) {
            const double u = warp_sum(sxx[q++]);
            if (lead && u != 0.0) {
                atomicAdd(&out[1 + NC + j * NC + k], u);
                if (k != j) atomicAdd(&out[1 + NC + k * NC + j], u);
            }
        }
    }
}

// ---- FP64 FMA pipe microbenchmark ------------------------------------------------------------
__global__ void trpl_dfma_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;   // never true; keeps the chain alive
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return TRPL_ECUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)

// pvSimPCR.py:327-331 (same expression order; pow() is glibc's, as in CPython)
void make_scales(double length, double time, int L, int T, double *sc, double *dx_o, double *dt_o)
{
    const double dx = length / L, dt = time / T;
    const double dx3 = pow(dx, 3.0);
    const double dtdx = dt / dx, dtdx2 = dtdx / dx, dtdx6 = dt / pow(dx, 6.0);
    sc[0] = dx3; sc[1] = dx3; sc[2] = dtdx2; sc[3] = dtdx2; sc[4] = dtdx2 / dx;
    sc[5] = dtdx; sc[6] = dtdx; sc[7] = dtdx6; sc[8] = dtdx6;
    sc[9] = 1.0 / dt; sc[10] = 1.0 / dt; sc[11] = 1.0 / dx;
    *dx_o = dx; *dt_o = dt;
}

struct Cfg { int M; bool pad; int W; };      // W = warps per simulation (1: warp kernel, >1: CTA kernel)
int pick_cfg(int L, Cfg *cfg)
{
    if (L < 2) return TRPL_EINVAL;
    if (L <= 256) {
        int M = 1;
        while (32 * M < L) M <<= 1;
        if (L % M == 0 && L >= 2 * M) {
            cfg->M = M; cfg->pad = (L != 32 * M); cfg->W = 1;
            return TRPL_OK;
        }
        if (L < 128) return TRPL_EUNSUPPORTED;    // odd small grids: need L % M == 0
    }
    // fine grids: 4 nodes per lane, W warps per simulation
    if (L % 4 != 0 || L > 128 * 16) return TRPL_EUNSUPPORTED;
    int W = 2;
    while (128 * W < L) W <<= 1;
    cfg->M = 4; cfg->pad = true; cfg->W = W;
    return TRPL_OK;
}

typedef void (*kern_t)(const KArgs);
kern_t pick_kernel(const Cfg &c)
{
    if (c.W > 1) {
        switch (c.W) {
        case 2: return trpl_sim_cta_kernel<2>;
        case 4: return trpl_sim_cta_kernel<4>;
        case 8: return trpl_sim_cta_kernel<8>;
        default: return trpl_sim_cta_kernel<16>;
        }
    }
    switch (c.M) {
    case 1: return c.pad ? trpl_sim_kernel<1, true> : trpl_sim_kernel<1, false>;
    case 2: return c.pad ? trpl_sim_kernel<2, true> : trpl_sim_kernel<2, false>;
    case 4: return c.pad ? trpl_sim_kernel<4, true> : trpl_sim_kernel<4, false>;
    default: return c.pad ? trpl_sim_kernel<8, true> : trpl_sim_kernel<8, false>;
    }
}

// threads per CTA, simulations per CTA and dynamic shared memory of a configuration
void cfg_shape(const Cfg &c, int *threads, int *sims_per_cta, size_t *smem)
{
    const size_t ring = (size_t)4 * 3 * c.M * 32 * sizeof(double);       // per warp
    if (c.W > 1) {
        *threads = c.W * 32;
        *sims_per_cta = 1;
        *smem = c.W * ring + (size_t)(2 * 3 * 32 * c.W + 2 * c.W * 4) * sizeof(double);
    } else {
        *threads = WARPS_PER_CTA * 32;
        *sims_per_cta = WARPS_PER_CTA;
        *smem = WARPS_PER_CTA * ring;
    }
}

int kernel_geometry(int device, const Cfg &cfg, kern_t *k_out, size_t *smem_out, int *ctas_per_sm,
                    int *sms)
{
    kern_t k = pick_kernel(cfg);
    int threads, spc; size_t smem;
    cfg_shape(cfg, &threads, &spc, &smem);
    CK(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((const void *)k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)k, threads, smem));
    if (nb < 1) return TRPL_EUNSUPPORTED;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    *k_out = k; *smem_out = smem; *ctas_per_sm = nb; *sms = nsm;
    return TRPL_OK;
}

int check_device(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return TRPL_ENODEVICE;
    }
    CK(cudaSetDevice(device));
    return TRPL_OK;
}

int launch_sims(KArgs &ka, const Cfg &cfg, int device, cudaStream_t st)
{
    kern_t k; size_t smem; int nb, nsm;
    int rc = kernel_geometry(device, cfg, &k, &smem, &nb, &nsm);
    if (rc) return rc;
    unsigned long long *counter = nullptr;
    CK(cudaMallocAsync((void **)&counter, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    ka.counter = counter;
    int threads, spc; size_t smem2;
    cfg_shape(cfg, &threads, &spc, &smem2);
    const unsigned long long items = (unsigned long long)ka.S * ka.C;
    unsigned long long want = (items + spc - 1) / spc;
    unsigned long long cap = (unsigned long long)nb * nsm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grid > 0) {
        k<<<grid, threads, smem, st>>>(ka);
        CK(cudaGetLastError());
    }
    CK(cudaFreeAsync(counter, st));
    return TRPL_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int trpl_version(void) { return 100; }

const char *trpl_error_string(int code)
{
    switch (code) {
    case TRPL_OK: return "ok";
    case TRPL_EINVAL: return "invalid argument";
    case TRPL_EUNSUPPORTED: return "unsupported shape (need L <= 256 with L = M*g, M in {1,2,4,8}, "
                                   "2 <= g <= 32, or L <= 2048 with L % 4 == 0; curves <= 8, "
                                   "observation files <= 4)";
    case TRPL_ECUDA: return "CUDA runtime error";
    case TRPL_ENODEVICE: return "no usable CUDA device";
    default: return "unknown error";
    }
}

const char *trpl_last_cuda_error(void) { return g_cuda_err; }

int trpl_resident_sims(int device, int L)
{
    Cfg cfg;
    int rc = pick_cfg(L, &cfg);
    if (rc) return rc;
    rc = check_device(device);
    if (rc) return rc;
    kern_t k; size_t smem; int nb, nsm;
    rc = kernel_geometry(device, cfg, &k, &smem, &nb, &nsm);
    if (rc) return rc;
    int threads, spc; size_t smem2;
    cfg_shape(cfg, &threads, &spc, &smem2);
    return nb * nsm * spc;
}

int trpl_solve_pl(const double *d_matpar, int64_t S, int64_t ld_matpar, const double *d_init,
                  double length, double time, int L, int T, int plT, int tol, int max_iter,
                  int max_order, int flags, void *d_pl, int pl_dtype, int64_t pl_stride,
                  int32_t *d_status, int64_t *d_iters, int device, void *stream)
{
    if (!d_matpar || !d_init || !d_pl || S < 0 || ld_matpar < TRPL_NPAR || T < 1 || plT < 1 ||
        max_iter < 1 || !(length > 0) || !(time > 0) || pl_stride < T / plT + 1 ||
        (pl_dtype != TRPL_F64 && pl_dtype != TRPL_F32))
        return TRPL_EINVAL;
    Cfg cfg;
    int rc = pick_cfg(L, &cfg);
    if (rc) return rc;
    rc = check_device(device);
    if (rc) return rc;
    if (S == 0) return TRPL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_order < 1 || max_order > 5) max_order = 5;

    KArgs ka;
    memset(&ka, 0, sizeof(ka));
    ka.x = d_matpar; ka.ldx = ld_matpar; ka.S = S; ka.pl_stride = pl_stride;
    ka.TOL = pow(10.0, -(double)tol);
    ka.sse = nullptr; ka.status = d_status; ka.iters = (long long *)d_iters;
    ka.mag_col = -1; ka.C = 1; ka.E = 0; ka.L = L; ka.plT = plT; ka.max_iter = max_iter;
    ka.max_order = max_order;
    ka.flags = (flags & TRPL_F_INIT_GRID_UNITS) | (pl_dtype == TRPL_F32 ? TRPL_F_EMULATE_F32 : 0);
    ka.pl_dtype = pl_dtype;
    CurveDev &cv = ka.curves[0];
    double dx, dt;
    make_scales(length, time, L, T, cv.scales, &dx, &dt);
    cv.init_mul = (flags & TRPL_F_INIT_GRID_UNITS) ? 1.0 : cv.scales[0];
    cv.redim = pow(dx, 2.0) * dt;
    cv.init = d_init;
    cv.pl_out = d_pl;
    cv.t_last = T;
    return launch_sims(ka, cfg, device, st);
}

int trpl_solve_loglik(const double *d_x, int64_t S, int64_t ldx, int mag_col,
                      const trpl_curve *curves, int C, int E, double time, int L, int T, int tol,
                      int max_iter, int max_order, int flags, double *d_sse, double *d_lnl,
                      int32_t *d_status, int64_t *d_iters, int device, void *stream)
{
    if (!d_x || !curves || !d_sse || !d_lnl || S < 0 || ldx < TRPL_NPAR || mag_col >= ldx ||
        T < 1 || max_iter < 1 || !(time > 0) || C < 1 || E < 1)
        return TRPL_EINVAL;
    if (C > TRPL_MAX_CURVES || E > TRPL_MAX_EXP) return TRPL_EUNSUPPORTED;
    Cfg cfg;
    int rc = pick_cfg(L, &cfg);
    if (rc) return rc;
    rc = check_device(device);
    if (rc) return rc;
    if (S == 0) return TRPL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_order < 1 || max_order > 5) max_order = 5;

    KArgs ka;
    memset(&ka, 0, sizeof(ka));
    ka.x = d_x; ka.ldx = ldx; ka.S = S; ka.pl_stride = 0;
    ka.TOL = pow(10.0, -(double)tol);
    ka.sse = d_sse; ka.iters = (long long *)d_iters;
    ka.mag_col = mag_col; ka.C = C; ka.E = E; ka.L = L; ka.plT = 1; ka.max_iter = max_iter;
    ka.max_order = max_order; ka.flags = flags; ka.pl_dtype = TRPL_F64;
    int *status_cs = nullptr;
    if (d_status) CK(cudaMallocAsync((void **)&status_cs, sizeof(int) * (size_t)C * S, st));
    ka.status = status_cs;
    for (int c = 0; c < C; c++) {
        const trpl_curve &src = curves[c];
        if (!src.d_init || !(src.length > 0)) return TRPL_EINVAL;
        CurveDev &cv = ka.curves[c];
        double dx, dt;
        make_scales(src.length, time, L, T, cv.scales, &dx, &dt);
        cv.init_mul = (flags & TRPL_F_INIT_GRID_UNITS) ? 1.0 : cv.scales[0];
        cv.redim = pow(dx, 2.0) * dt;
        cv.init = src.d_init;
        cv.pl_out = nullptr;
        cv.t_last = 0;
        for (int e = 0; e < E; e++) {
            const trpl_obs &o = src.obs[e];
            if (o.n < 0 || (o.n > 0 && (!o.d_hi || !o.d_whi || !o.d_wlo || !o.d_val)))
                return TRPL_EINVAL;
            if (o.n > 0 && (o.hi_max < 1 || o.hi_max > T)) return TRPL_EINVAL;
            cv.obs[e].n = o.n; cv.obs[e].hi = o.d_hi; cv.obs[e].whi = o.d_whi;
            cv.obs[e].wlo = o.d_wlo; cv.obs[e].val = o.d_val;
            if (o.n > 0 && o.hi_max > cv.t_last) cv.t_last = o.hi_max;   // causal truncation
        }
    }
    rc = launch_sims(ka, cfg, device, st);
    if (rc) return rc;
    const int tb = 256;
    trpl_finish_kernel<<<(unsigned)((S + tb - 1) / tb), tb, 0, st>>>(d_sse, d_lnl, status_cs, d_status,
                                                                      S, C, E);
    CK(cudaGetLastError());
    if (status_cs) CK(cudaFreeAsync(status_cs, st));
    return TRPL_OK;
}

int trpl_log10_clamp(void *d_pl, int dtype, int64_t n, double min, int device, void *stream)
{
    if (!d_pl || n < 0 || (dtype != TRPL_F64 && dtype != TRPL_F32)) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) return TRPL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int tb = 256;
    long long blocks = (n + tb - 1) / tb;
    if (blocks > (long long)nsm * 16) blocks = (long long)nsm * 16;
    if (dtype == TRPL_F64)
        trpl_log10_kernel_f64<<<(unsigned)blocks, tb, 0, st>>>((double *)d_pl, n, min);
    else
        trpl_log10_kernel_f32<<<(unsigned)blocks, tb, 0, st>>>((float *)d_pl, n, min);
    CK(cudaGetLastError());
    return TRPL_OK;
}

int trpl_lnp_accumulate(double *d_P, const double *d_pl, int64_t S, int64_t n, int64_t ld,
                        const double *d_values, const double *d_mag, int device, void *stream)
{
    if (!d_P || !d_pl || !d_values || !d_mag || S < 0 || n < 0 || ld < n) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    if (S == 0) return TRPL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int tb = 256;
    long long blocks = (S * 32 + tb - 1) / tb;
    if (blocks > (long long)nsm * 8) blocks = (long long)nsm * 8;
    trpl_lnp_kernel<<<(unsigned)blocks, tb, 0, st>>>(d_P, d_pl, S, n, ld, d_values, d_mag);
    CK(cudaGetLastError());
    return TRPL_OK;
}

int trpl_obs_prepare(const double *times, int32_t n, double time, int T, int32_t *hi, double *whi,
                     double *wlo)
{
    if (!times || !hi || !whi || !wlo || n < 0 || T < 1 || !(time > 0)) return TRPL_EINVAL;
    // numpy.linspace(0, time, T+1): x_i = i*step with step = time/T, and x_T = time exactly
    const double step = time / T;
    int maxhi = 0;
    double prev = -INFINITY;
    for (int32_t i = 0; i < n; i++) {
        const double t = times[i];
        if (!(t >= 0.0) || !(t <= time) || t < prev) return TRPL_EINVAL;
        prev = t;
        // searchsorted(side='left'): first index k with x_k >= t
        long long k = (long long)floor(t / step);
        if (k < 0) k = 0;
        if (k > T) k = T;
        auto xk = [&](long long q) { return q >= T ? time : (double)q * step; };
        while (k > 0 && xk(k - 1) >= t) k--;
        while (k < T && xk(k) < t) k++;
        if (k < 1) k = 1;
        if (k > T) k = T;
        hi[i] = (int32_t)k;
        const double x_lo = xk(k - 1), x_hi = xk(k);
        whi[i] = (t - x_lo) / (x_hi - x_lo);
        wlo[i] = (x_hi - t) / (x_hi - x_lo);
        if (k > maxhi) maxhi = (int)k;
    }
    return maxhi;
}

int trpl_lse_partial(const double *d_x, int64_t n, double *d_out2, int device, void *stream)
{
    if (!d_x || !d_out2 || n < 0) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    trpl_lse_init_kernel<<<1, 1, 0, st>>>(d_out2);
    if (n > 0) {
        int nsm = 0;
        CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
        const int tb = 256;
        long long blocks = (n + tb - 1) / tb;
        if (blocks > (long long)nsm * 8) blocks = (long long)nsm * 8;
        trpl_lse_max_kernel<<<(unsigned)blocks, tb, 0, st>>>(d_x, n, d_out2);
        trpl_lse_sum_kernel<<<(unsigned)blocks, tb, 0, st>>>(d_x, n, d_out2);
    }
    CK(cudaGetLastError());
    return TRPL_OK;
}


int trpl_random_grid(double *d_x, int64_t S, int64_t ldx, const double *minx, const double *maxx,
                     const int32_t *do_log, int ncol, int override_flags, uint64_t seed,
                     uint64_t first_sample, int device, void *stream)
{
    if (!d_x || !minx || !maxx || !do_log || S < 0 || ncol < 1 || ncol > 16 || ldx < ncol) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    if (S == 0) return TRPL_OK;
    GridArgs ga;
    memset(&ga, 0, sizeof(ga));
    for (int j = 0; j < ncol; j++) {
        if (!(minx[j] <= maxx[j]) || (do_log[j] && minx[j] != maxx[j] && !(minx[j] > 0))) return TRPL_EINVAL;
        ga.lo[j] = minx[j]; ga.hi[j] = maxx[j]; ga.do_log[j] = do_log[j];
    }
    ga.ncol = ncol;
    ga.eq_mu = (override_flags & 1) && ncol > 3;
    ga.eq_s = (override_flags & 2) && ncol > 6;
    ga.eq_auger = (override_flags & 4) && ncol > 8;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int tb = 256;
    long long blocks = (S * ncol + tb - 1) / tb;
    if (blocks > (long long)nsm * 16) blocks = (long long)nsm * 16;
    trpl_random_grid_kernel<<<(unsigned)blocks, tb, 0, (cudaStream_t)stream>>>(d_x, S, ldx, ga, seed, first_sample);
    CK(cudaGetLastError());
    return TRPL_OK;
}

int trpl_posterior_weights(const double *d_lnp, int64_t n, double lse, double *d_w, int device, void *stream)
{
    if (!d_lnp || !d_w || n < 0) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) return TRPL_OK;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int tb = 256;
    long long blocks = (n + tb - 1) / tb;
    if (blocks > (long long)nsm * 16) blocks = (long long)nsm * 16;
    trpl_weights_kernel<<<(unsigned)blocks, tb, 0, (cudaStream_t)stream>>>(d_lnp, n, lse, d_w);
    CK(cudaGetLastError());
    return TRPL_OK;
}

int trpl_weighted_hist(const double *d_x, int64_t n, int64_t ldx, int colx, int coly, const double *d_w,
                       double lox, double hix, int nbx, double loy, double hiy, int nby, double *d_hist,
                       int device, void *stream)
{
    if (!d_x || !d_hist || n < 0 || colx < 0 || colx >= ldx || coly >= ldx || nbx < 1 || !(hix > lox))
        return TRPL_EINVAL;
    if (coly >= 0 && (nby < 1 || !(hiy > loy))) return TRPL_EINVAL;
    const long long nb = (long long)nbx * (coly >= 0 ? nby : 1);
    if (nb > 8192) return TRPL_EUNSUPPORTED;
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) return TRPL_OK;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int tb = 256;
    long long blocks = (n + tb - 1) / tb;
    if (blocks > (long long)nsm * 4) blocks = (long long)nsm * 4;
    const size_t smem = (size_t)nb * sizeof(double);
    CK(cudaFuncSetAttribute((const void *)trpl_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    trpl_hist_kernel<<<(unsigned)blocks, tb, smem, (cudaStream_t)stream>>>(d_x, ldx, colx, coly, d_w, n, lox, hix,
                                                                            nbx, loy, hiy, nby, d_hist);
    CK(cudaGetLastError());
    return TRPL_OK;
}

int trpl_weighted_moments(const double *d_x, int64_t n, int64_t ldx, int ncol, const double *d_w,
                          double *d_out, int device, void *stream)
{
    if (!d_x || !d_w || !d_out || n < 0 || ncol < 1 || ncol > 15 || ldx < ncol) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) return TRPL_OK;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    if (ncol == 13 || ncol == 12) {
        long long nb = (n + 127) / 128;
        if (nb > (long long)nsm * 2) nb = (long long)nsm * 2;
        if (ncol == 13)
            trpl_moments_kernel_fixed<13><<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(d_x, ldx, d_w, n, d_out);
        else
            trpl_moments_kernel_fixed<12><<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(d_x, ldx, d_w, n, d_out);
        CK(cudaGetLastError());
        return TRPL_OK;
    }
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)nsm * 8) blocks = (long long)nsm * 8;
    const size_t smem = (size_t)(1 + ncol + ncol * ncol) * sizeof(double);
    trpl_moments_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(d_x, ldx, ncol, d_w, n, d_out);
    CK(cudaGetLastError());
    return TRPL_OK;
}

int trpl_bench_dfma(int device, int iters, double *tflops, double *ms)
{
    if (!tflops || iters < 1) return TRPL_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    double *out = nullptr;
    CK(cudaMalloc((void **)&out, sizeof(double)));
    const int tb = 256, blocks = nsm * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    trpl_dfma_kernel<<<blocks, tb>>>(out, iters / 4 + 1, 1.0);   // warm-up
    CK(cudaEventRecord(e0));
    trpl_dfma_kernel<<<blocks, tb>>>(out, iters, 1.0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    const double flop = 2.0 * 8 * 16 * (double)iters * tb * (double)blocks;
    *tflops = flop / (t * 1e-3) / 1e12;
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return TRPL_OK;
}

}  // extern "C"

/*
 * trpl_b200.h -- C ABI of libtrpl_b200.so: the B200 (sm_100a) engine for the TRPL
 * forward-model + likelihood hot path of HagesLab/Bayesian-Inference-TRPL.
 *
 * Every entry point takes plain pointers and sizes.  Pointers named d_* are DEVICE
 * pointers (e.g. torch.Tensor.data_ptr()); everything else is host memory that is
 * only read during the call.  All calls are asynchronous on `stream` (a cudaStream_t
 * passed as void*, NULL = legacy default stream) of CUDA device `device`; the caller
 * synchronises.  The library keeps no mutable global state and is re-entrant per
 * device/stream (the reference's module globals SIZ/BuSIZ/MSPB, pvSimPCR.py:319-324,
 * have no counterpart).
 *
 * Return value: 0 on success, a negative TRPL_E* code otherwise
 * (trpl_error_string() maps it to text).  Solver failures are NOT call failures: they
 * are reported per sample in `d_status` (the reference aborts the whole launch and
 * leaves garbage, pvSimPCR.py:269-274,290-292).
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   trpl_solve_pl        pvSimPCR.pvSim -> tEvol -> iterate/pcreduce/norm2   pvSimPCR.py:14-401
 *   trpl_log10_clamp     probs.fastlog  -> log_kernel                        probs.py:64-85
 *   trpl_lnp_accumulate  probs.prob     -> kernel_lnP                        probs.py:20-62
 *   trpl_solve_loglik    the whole per-sample pipeline bayeslib.simulate drives
 *                        (model -> self_normalize -> fastlog -> griddata -> prob,
 *                        summed over curves)                                 bayeslib.py:117-201
 *   trpl_obs_prepare     scipy griddata/interp1d index+weight rule used at   bayeslib.py:186-189
 *   trpl_lse_partial     Visualization/utils.normalize (shifted exp / sum)   Visualization/utils.py:157-166
 *   trpl_posterior_weights, trpl_weighted_hist, trpl_weighted_moments
 *                        normalize / marginalize_1D / marginalize_2D / w_mean /
 *                        w_variance / covariance                             Visualization/utils.py:157-285
 *   trpl_random_grid     bayeslib.random_grid + make_grid overrides          bayeslib.py:18-76
 */
#ifndef TRPL_B200_H
#define TRPL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRPL_NPAR        12   /* n0,p0,DN,DP,B,Sf,Sb,CN,CP,tauN,tauP,Lambda  (pvSimPCR.py:97-108) */
#define TRPL_MAX_CURVES   8   /* excitation curves per fused call            */
#define TRPL_MAX_EXP      4   /* observation files per fused call            */

/* error codes */
#define TRPL_OK            0
#define TRPL_EINVAL       -1  /* bad argument                                */
#define TRPL_EUNSUPPORTED -2  /* shape outside what the kernels cover        */
#define TRPL_ECUDA        -3  /* CUDA runtime error (see trpl_last_cuda_error) */
#define TRPL_ENODEVICE    -4  /* no CUDA device / device is not sm_100       */

/* dtype codes for PL buffers */
#define TRPL_F64 0
#define TRPL_F32 1

/* per-sample status bits */
#define TRPL_ST_NOCONV    1   /* Newton loop hit max_iter (pvSimPCR.py:269)  */
#define TRPL_ST_NONFINITE 2   /* residual became NaN/Inf                     */

/* flags of trpl_solve_pl / trpl_solve_loglik */
#define TRPL_F_INIT_GRID_UNITS 1   /* init profile already multiplied by dx^3 ("exp" mode, pvSimPCR.py:347-353) */
#define TRPL_F_LOG_PL          2   /* compare log10(PL)      (sim_flags["log_pl"],         bayeslib.py:155)    */
#define TRPL_F_SELF_NORMALIZE  4   /* PL /= PL[t=0]          (sim_flags["self_normalize"], bayeslib.py:150)    */
#define TRPL_F_EMULATE_F32     8   /* reproduce the float32 PL buffer of bayeslib.py:137 (store, /=, log10f)   */

/* One observation set (one curve of one observation file), device arrays of length n.
 * Produced by trpl_obs_prepare + upload.  Observation i is compared with
 *   y = whi*l[hi] + wlo*l[hi-1]                    (l = log10 PL on the step grid),
 * the linear rule scipy.interpolate.interp1d applies for bayeslib.py:186-189.            */
typedef struct trpl_obs {
    int32_t        n;
    int32_t        hi_max;  /* largest d_hi entry (return value of trpl_obs_prepare)              */
    const int32_t *d_hi;    /* [n] upper bracketing step index, 1..T, non-decreasing      */
    const double  *d_whi;   /* [n] (t_obs - t[hi-1]) / (t[hi] - t[hi-1])                  */
    const double  *d_wlo;   /* [n] (t[hi] - t_obs)   / (t[hi] - t[hi-1])                  */
    const double  *d_val;   /* [n] observed log10 PL (or PL when !LOG_PL)                 */
} trpl_obs;

/* One excitation curve. */
typedef struct trpl_curve {
    const double *d_init;   /* [L] initial excess carrier density, nm^-3 (iniPar row, bayes_io.py:106-119) */
    double        length;   /* film thickness for this curve, nm (bayeslib.py:109-119)    */
    trpl_obs      obs[TRPL_MAX_EXP];
} trpl_curve;

int         trpl_version(void);
const char *trpl_error_string(int code);
const char *trpl_last_cuda_error(void);          /* thread-local text of the last CUDA failure */

/* Number of work items (sample x curve simulations) resident at once on `device` with the
 * kernel configuration used for L nodes; bench/tests use it to size batches in whole waves. */
int trpl_resident_sims(int device, int L);

/*
 * Forward model for ONE curve: PL(t) of S parameter samples.   [pvSimPCR.pvSim]
 *   d_matpar  [S][ld_matpar] physical units (nm, ns, V); first 12 columns are used
 *   d_init    [L]
 *   L,T,plT,tol,max_iter  simPar = [Length,Time,L,T,plT,pT,tol,MAX] (parallel_bayes_gpu.py:81)
 *   max_order BDF order cap, 5 = reference (pvSimPCR.py:241-250); 2 = Legacy/pvSim.py:102-105
 *   d_pl      [S][pl_stride] PL in nm^-2 ns^-1, T/plT+1 values per row, dtype TRPL_F64|TRPL_F32
 *             (F32: value rounded on store, then divided by dx^2*dt in float32, pvSimPCR.py:384,393)
 *   d_status  [S] int32 or NULL;  d_iters [S] int64 total Newton iterations or NULL
 */
int trpl_solve_pl(const double *d_matpar, int64_t S, int64_t ld_matpar, const double *d_init,
                  double length, double time, int L, int T, int plT, int tol, int max_iter,
                  int max_order, int flags, void *d_pl, int pl_dtype, int64_t pl_stride,
                  int32_t *d_status, int64_t *d_iters, int device, void *stream);

/*
 * Fused forward model + likelihood for C curves and E observation files.
 *   d_x      [S][ldx]   sample matrix; columns 0..11 = matPar, `mag_col` = mag_offset (12 in X[S,13])
 *   curves   host array [C]
 *   d_sse    [E][C][S]  scratch: per-curve sums of squared residuals (written)
 *   d_lnl    [E][S]     in/out: lnL[e][s] -= sum_c sse[e][c][s], curves in order (probs.py:60, bayeslib.py:195)
 *   d_status [S] int32 (OR over curves) or NULL;  d_iters [C][S] int64 or NULL
 * The time integration of curve c stops at the last step any of its observations needs
 * (causal, identical output).
 */
int trpl_solve_loglik(const double *d_x, int64_t S, int64_t ldx, int mag_col,
                      const trpl_curve *curves, int C, int E, double time, int L, int T, int tol,
                      int max_iter, int max_order, int flags, double *d_sse, double *d_lnl,
                      int32_t *d_status, int64_t *d_iters, int device, void *stream);

/* In-place log10(max(x, min)) with the reference's float32/float64 semantics.   [probs.fastlog] */
int trpl_log10_clamp(void *d_pl, int dtype, int64_t n, double min, int device, void *stream);

/* d_P[j] -= sum_i (d_pl[j*ld+i] + d_mag[j] - d_values[i])^2.                     [probs.prob]    */
int trpl_lnp_accumulate(double *d_P, const double *d_pl, int64_t S, int64_t n, int64_t ld,
                        const double *d_values, const double *d_mag, int device, void *stream);

/* HOST helper: bracket n observation times on the step grid t_i = i*(time/T) (numpy.linspace(0,
 * time, T+1)) with scipy interp1d's rule (hi = clip(searchsorted_left, 1, T)).  Returns the
 * largest hi, or TRPL_EINVAL if a time lies outside [0, time] or times are not sorted.       */
int trpl_obs_prepare(const double *times, int32_t n, double time, int T,
                     int32_t *hi, double *whi, double *wlo);

/* Shard-local part of the posterior normalisation: d_out[0] = max_i x_i, d_out[1] =
 * sum_i exp(x_i - max) over finite x (NaN skipped like numpy.nanmax/nansum).               */
int trpl_lse_partial(const double *d_x, int64_t n, double *d_out2, int device, void *stream);

/* Sample matrix on the device: X[s][j] uniform (or log-uniform where do_log[j]) in [minx[j], maxx[j]],
 * constant where minx[j] == maxx[j]; counter-based Philox4x32-10 keyed by `seed`, sample index
 * first_sample + s (so shards of one global draw can be generated independently on every rank).
 * override_flags: bit0 X[:,2]=X[:,3] (equal mu), bit1 X[:,6]=X[:,5] (equal S), bit2 X[:,8]=X[:,7]
 * (equal Auger), bayeslib.py:68-75.  minx/maxx/do_log are HOST arrays of ncol (<= 16) entries.   */
int trpl_random_grid(double *d_x, int64_t S, int64_t ldx, const double *minx, const double *maxx,
                     const int32_t *do_log, int ncol, int override_flags, uint64_t seed,
                     uint64_t first_sample, int device, void *stream);

/* d_w[i] = exp(d_lnp[i] - lse)  (NaN -> 0): posterior weights once lse = log sum exp is known.   */
int trpl_posterior_weights(const double *d_lnp, int64_t n, double lse, double *d_w, int device,
                           void *stream);

/* Weighted histogram with numpy.histogram(2d) binning: nbx uniform bins on [lox, hix] of column
 * colx (and nby bins on [loy, hiy] of column coly when coly >= 0; row-major [nbx][nby]).
 * d_hist is ACCUMULATED into (zero it first); d_w == NULL counts samples.  Raw sums, no density
 * normalisation (marginalize_1D/2D divide by sum*bin width on the host).                         */
int trpl_weighted_hist(const double *d_x, int64_t n, int64_t ldx, int colx, int coly, const double *d_w,
                       double lox, double hix, int nbx, double loy, double hiy, int nby, double *d_hist,
                       int device, void *stream);

/* Raw weighted moments, ACCUMULATED into d_out[1 + ncol + ncol*ncol]:
 * [0] = sum w, [1+j] = sum w x_j, [1+ncol+j*ncol+k] = sum w x_j x_k  (ncol <= 15).               */
int trpl_weighted_moments(const double *d_x, int64_t n, int64_t ldx, int ncol, const double *d_w,
                          double *d_out, int device, void *stream);

/* Self-test hook: y[i] = the solver's reciprocal of x[i] (MUFU.RCP64H seed + one cubically convergent
 * step; replaces the reference's FP64 divisions, pvSimPCR.py:57-69,156-209).  Domain = what the solver
 * feeds it: finite, normal, |x| in [2^-1000, 2^1000]; max error 1 ulp there.  Zero, denormal, infinite or
 * NaN operands give a non-finite result, which the solver reports as TRPL_ST_NONFINITE for that sample
 * (the reference would divide by zero at the same place).                                        */
int trpl_selftest_rcp(const double *d_x, double *d_y, int64_t n, int device, void *stream);

/* FP64 FMA-pipe microbenchmark (roofline denominator): runs `iters` dependent-chain DFMA
 * rounds on every SM and returns the achieved TFLOP/s (2 flop per FMA) in *tflops.
 * Synchronous.                                                                             */
int trpl_bench_dfma(int device, int iters, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* TRPL_B200_H */

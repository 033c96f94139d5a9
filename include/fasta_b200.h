/*
 * fasta_b200.h -- C ABI of libfasta_b200.so: the sm_100a kernels under FASTA's
 * forward-backward-splitting loop.
 *
 * The reference (phasepack/fasta-python) has no FFI: its "operator interface" is the Python
 * callable protocol of fasta.fasta() (reference fasta/__init__.py:38-53).  Each entry point below
 * replaces the numpy expression(s) cited next to it; INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add to route those expressions here.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure (fb200_last_error() has text);
 *     nothing throws across the boundary;
 *   - all pointers are BORROWED device pointers (fp64, densely packed unless an lda is given);
 *     the library allocates nothing persistent;
 *   - `ws` is a caller-owned, zero-initialised device workspace of at least
 *     fb200_workspace_bytes(M, N) bytes, reused by consecutive calls on one stream;
 *   - `scal` is a caller-owned device array of FB200_NSCAL doubles that receives the reduction
 *     results (slot indices FB200_S_*); the caller copies it to the host once per decision point;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise;
 *   - reductions are atomics-free and fixed-order: results are bit-reproducible run to run.
 */
#ifndef FASTA_B200_H
#define FASTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_ABI_VERSION 1
#define FB200_NSCAL 64
#define FB200_FRING 40

/* slots of the device scalar block */
enum {
    FB200_S_F        = 0,  /* raw loss reduction: sum (z-b)^2 (least squares) or the logistic sum */
    FB200_S_DX_G0    = 1,  /* <Dx, gradf0>                      __init__.py:200 */
    FB200_S_DX_SQ    = 2,  /* <Dx, Dx>                          __init__.py:200,258,272 */
    FB200_S_XMXH_SQ  = 3,  /* <x1 - x1hat, x1 - x1hat>          __init__.py:274 */
    FB200_S_PEN      = 4,  /* raw penalty reduction: sum |x1|   sparse_least_squares.py:43 */
    FB200_S_RESTART  = 5,  /* <x0 - x1, x1 - x_accel0>          __init__.py:231 */
    FB200_S_DX_DG    = 6,  /* <Dx, Dg>                          __init__.py:255 */
    FB200_S_DG_SQ    = 7,  /* <Dg, Dg>                          __init__.py:260 */
    FB200_S_G1_SQ    = 8,  /* <gradf1, gradf1> (next iteration's |gradf0|)  __init__.py:274 */
    FB200_S_THETA    = 9,  /* l1-ball projection threshold      proximal.py:26 */
    FB200_S_AUX0     = 10, /* dot / norm helpers write here by default */
    FB200_S_AUX1     = 11,
    FB200_S_AUX2     = 12,
    FB200_S_AUX3     = 13,
    FB200_S_TAU      = 14, /* step size for the NEXT trial, written by fb200_trial_decide (__init__.py:253-270) */
    FB200_S_TAU_USED = 15, /* step size the last trial ran with */
    FB200_S_SKIP     = 16, /* != 0: trials queued with tau = NaN return at once (the trial before them was rejected by
                              the line search, or the stop rule fired) */
    FB200_S_SKIPPED  = 17, /* != 0: THIS trial was skipped (its other slots are stale) */
    FB200_S_IT       = 18, /* accepted trials so far (the loop index i of __init__.py:172) */
    FB200_S_MAXRES   = 19, /* running maximum of the residuals (__init__.py:281) */
    FB200_S_G0SQ     = 20, /* |gradf0|^2 of the next trial = |gradf1|^2 of the last accepted one (__init__.py:274) */
    FB200_S_FRING    = 24  /* ring of the last FB200_FRING values of f_hist, index k at slot FB200_S_FRING + k % FB200_FRING */
};

/* loss tags: f and gradf evaluated on z = A x */
enum {
    FB200_LOSS_NONE          = 0,  /* no loss: only z is produced */
    FB200_LOSS_LEAST_SQUARES = 1,  /* f=.5|z-b|^2, gradf=z-b        sparse_least_squares.py:41-42 */
    FB200_LOSS_LOGISTIC      = 2   /* f=sum log(1+e^z)-(b==1)z, gradf=-b/(1+e^{bz})  sparse_logistic.py:47-48 */
};

/* prox tags: x1 = prox(x1hat, t) */
enum {
    FB200_PROX_IDENTITY = 0,  /* g is None                          __init__.py:88-90 */
    FB200_PROX_SHRINK   = 1,  /* sign(x)*max(|x|-p0,0)              proximal.py:67 */
    FB200_PROX_NONNEG   = 2,  /* max(x,0)                           nn_least_squares.py:42 */
    FB200_PROX_BOX      = 3,  /* min(max(x,p0),p1)                  svm.py:71 */
    FB200_PROX_L1BALL   = 4,  /* shrink by scal[FB200_S_THETA]      proximal.py:34-41 (after fb200_l1ball_threshold) */
    FB200_PROX_TV_BALL  = 5   /* pairs (y0,y1) / max(|y|_2, 1)      tv_denoising.py:89-96 */
};

int         fb200_abi_version(void);
const char* fb200_last_error(void);
/* bytes of zero-initialised workspace needed for an M x N dense map (or vectors up to max(M,N)) */
size_t      fb200_workspace_bytes(int64_t M, int64_t N);
/* 1 if the TMA (cp.async.bulk.tensor) streaming path will be used for this matrix, else 0 */
int         fb200_dense_uses_tma(const double* A, int64_t lda, int64_t M, int64_t N);

/* ---- K1-K3: forward step, backward (prox) step and their reductions -------------------------
 * xhat = x0 - tau*g0 ; x1 = prox(xhat) ; dx = x1 - x0            __init__.py:181,184,186 (redo :207-211)
 * scal[S_DX_G0, S_DX_SQ, S_XMXH_SQ, S_PEN] and, when xa_prev != NULL, scal[S_RESTART].       */
int fb200_fbs_step(const double* x0, const double* g0, double tau, int prox, double p0, double p1,
                   const double* xa_prev, int64_t n, double* xhat, double* x1, double* dx,
                   double* scal, void* ws, void* stream);

/* generic-path pieces of the same step, for untagged (user-callable) prox operators:
 *   forward_step: xhat = x0 - tau*g0                              __init__.py:181
 *   step_reduce : dx = x1 - x0 and the reductions of fbs_step     __init__.py:186,200,231,274
 *   prox_apply  : out = prox(x) for a tagged prox (stand-alone use of fasta.proximal.*)       */
int fb200_forward_step(const double* x0, const double* g0, double tau, int64_t n, double* xhat,
                       void* stream);
int fb200_step_reduce(const double* x0, const double* x1, const double* xhat, const double* g0,
                      const double* xa_prev, int64_t n, double* dx, double* scal, void* ws,
                      void* stream);
int fb200_prox_apply(const double* x, int prox, double p0, double p1, int64_t n, double* out,
                     const double* scal, void* stream);

/* threshold of the Euclidean projection of v onto {|x|_1 <= radius} -> scal[S_THETA]
 * (0 when v is already inside)                                    proximal.py:12-41            */
int fb200_l1ball_threshold(const double* v, int64_t n, double radius, double* scal, void* ws,
                           void* stream);

/* ---- K9: FISTA extrapolation                                     __init__.py:242-245,274,285
 * x1 = xa1 + c*(xa1 - xa0) ; z1 = za1 + c*(za1 - za0) ; r = gradf(z1)
 * scal[S_F] (loss at z1), scal[S_XMXH_SQ] (|x1 - xhat|^2), scal[S_PEN] (penalty of x1)        */
int fb200_accel_step(double c, const double* xa1, const double* xa0, const double* xhat, int64_t n,
                     double* x1, const double* za1, const double* za0, const double* b, int64_t m,
                     int loss, int prox, double* z1, double* r, double* scal, void* ws, void* stream);

/* ---- K5 stand-alone: r = gradf(z), scal[S_F] = raw f(z)          sparse_least_squares.py:41-42 */
int fb200_loss_eval(int loss, const double* z, const double* b, int64_t m, double* r, double* scal,
                    void* ws, void* stream);

/* ---- K8: Barzilai-Borwein reductions                             __init__.py:254-260,274
 * dg = g1 + (xhat - x0)/tau ; scal[S_DX_DG, S_DG_SQ] (if adaptive) and scal[S_G1_SQ]         */
int fb200_bb_reduce(const double* g1, const double* x0, const double* xhat, const double* dx,
                    double tau, int64_t n, int adaptive, double* scal, void* ws, void* stream);

/* ---- K4+K5: z = A x fused with the loss epilogue                 linalg.py:41 (A @ x), __init__.py:187-188
 * A is M x N row-major with leading dimension lda (elements).  z, r are length M.
 * loss == FB200_LOSS_NONE: only z is written (r, b may be NULL).                              */
int fb200_gemv_loss(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss,
                    const double* b, double* z, double* r, double* scal, void* ws, size_t ws_bytes,
                    void* stream);

/* ---- K7+K8: g = A^T r fused with the Barzilai-Borwein epilogue   linalg.py:41 (A.T @ x), __init__.py:248-260
 * bb: 0 = only g; 1 = g and scal[S_G1_SQ]; 2 = also scal[S_DX_DG, S_DG_SQ] (needs x0,xhat,dx,tau) */
int fb200_gemvT_bb(const double* A, int64_t lda, int64_t M, int64_t N, const double* r, double* g,
                   int bb, const double* x0, const double* xhat, const double* dx, double tau,
                   double* scal, void* ws, size_t ws_bytes, void* stream);

/* ---- K4+K5+K7+K8 in ONE pass over A (single-pass sweep; csrc/dense_sweep.cu) ---------------------
 * z = A x, r = gradf(z), scal[S_F], and g = A^T r with the BB epilogue (as gemv_loss followed by
 * gemvT_bb, reference __init__.py:187-188 and :248-260) while streaming A from HBM once: a row of A
 * stays resident in the distributed shared memory of a thread-block cluster between the two uses.
 * g == NULL leaves only the per-cluster partials (multi-GPU callers all-reduce first).
 * fb200_sweep_supported returns the cluster size that will be used, or 0 if the matrix is not
 * eligible (then use the two-pass entry points).                                               */
int fb200_sweep_supported(const double* A, int64_t lda, int64_t M, int64_t N);
/* plan[0..4] = cluster size, co-resident clusters, pipeline stages, columns per CTA, double2 per thread */
int fb200_sweep_plan(int64_t M, int64_t N, int* plan);
int fb200_dense_sweep(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss,
                      const double* b, double* z, double* r, double* g, int bb, const double* x0,
                      const double* xhat, const double* dx, double tau, double* scal, void* ws,
                      size_t ws_bytes, void* stream);

/* the same single pass in FISTA (accelerated) mode, reference __init__.py:187-188,242-249: xa1 is the prox point,
 * za1 = A xa1 its image (written), the gradient is taken at the extrapolated z = za1 + c (za1 - za0) (written to z,
 * r = gradf(z)); scal[S_F] = f(za1) for the line search, scal[S_AUX3] = f(z); g and the BB epilogue as above   */
int fb200_dense_sweep_accel(const double* A, int64_t lda, int64_t M, int64_t N, const double* xa1, int loss,
                            const double* b, const double* za0, double c, double* za1, double* z, double* r,
                            double* g, int bb, const double* x0, const double* xhat, const double* dx, double tau,
                            double* scal, void* ws, size_t ws_bytes, void* stream);

/* ---- device-resident FBS loop for small dense problems (csrc/resident_loop.cu) -------------------------
 * The whole loop of reference __init__.py:172-313 (plain, adaptive and FISTA modes; accelerate != 0 needs the prox
 * point / image ping-pong buffers xa_*, za_* with xa_a = start point, za_a = A x0, and alpha_h; bit 30 of bt_h[i]
 * flags a restart of the acceleration in iteration i; built-in stop rules 0..3 =
 * stopping.residual / norm_residual / ratio_residual / hybrid_residual, elementwise prox) in ONE cooperative
 * kernel: no host round trip per iteration.  x_a / g_a hold the start point and its gradient, f_h[0] (and
 * obj_h[0]) the start values; histories, per-iteration backtrack counts and %globaltimer stamps are device arrays
 * of max_iters (+1) entries; out[0..3] = iterations, total backtracks, index of the buffer with the last iterate,
 * 1 if the single-cluster variant ran;
 * best receives the best iterate.  fb200_resident_blocks returns 0 when the problem is not eligible (A must fit
 * the L2); part: fb200_resident_scratch_doubles(M, N) doubles.                                              */
int fb200_resident_blocks(int64_t M, int64_t N);
size_t fb200_resident_scratch_doubles(int64_t M, int64_t N);
/* 1 if fb200_resident_fbs will run its single-cluster variant for an M x N problem: one 16-CTA thread-block cluster
 * keeps the matrix in its shared memory (row bands for A x, column groups for A^T r), phases separated by the hardware
 * cluster barrier; out[3] of fb200_resident_fbs reports which variant ran.  FASTA_B200_RESIDENT_CLUSTER=0 disables it. */
int fb200_resident_cluster_ok(int64_t M, int64_t N);
int fb200_resident_fbs(const double* A, int64_t lda, int64_t M, int64_t N, const double* b, int loss, int prox,
                       double pen_mu, double p_lo, double p_hi, double* x_a, double* x_b, double* g_a, double* g_b,
                       double* xhat, double* dx, double* best, double* z, double* r, double* part,
                       double* resid_h, double* nresid_h, double* tau_h, double* f_h, double* obj_h, int* bt_h,
                       unsigned long long* clock_h, double* out, double tau_init, double g1_sq_init,
                       double tolerance, double shrink, int adaptive, int backtrack, int window,
                       int max_backtracks, int max_iters, int stop_rule, int evaluate_objective, int accelerate,
                       int restart, double* xa_a, double* xa_b, double* za_a, double* za_b, double* alpha_h,
                       void* stream);

/* ---- K14: batched contractions (B columns, batch index fastest), fp64 DMMA GEMM -----------------
 * adjoint = 0:  C (Mg x Ng) = A (Mg x K) . B (K x Ng)         the per-column `A @ x`   of linalg.py:41
 * adjoint = 1:  C (Mg x Ng) = A^T . B with A stored (K x Mg)  the per-column `A.T @ r` of linalg.py:41
 * all matrices row-major fp64 with even leading dimensions and 16-byte aligned bases.            */
int fb200_gemm_f64(int adjoint, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, int64_t Mg, int64_t Ng, int64_t K, int splits, int64_t split_stride,
                   void* stream);
/* K slices that balance the chip for this shape; with splits > 1 C holds [splits][split_stride]
 * partial products which the batched epilogues add in index order                                */
int fb200_gemm_splits(int64_t Mg, int64_t Ng, int64_t K);

/* ---- K14 on tcgen05: fp64-accurate batched contractions from int8 digit planes (csrc/ozaki_gemm.cu) --
 * Both operands are split error-free into 8 signed 7-bit digit planes with a power-of-two scale per row
 * (left operand) / per column (right operand); the 36 digit-pair products with s + t <= 7 run as
 * tcgen05.mma.kind::i8 GEMMs into int32 TMEM accumulators and are recombined in fp64.  Replaces the same
 * reference lines as fb200_gemm_f64 (linalg.py:41, per column).
 *   pad(n, tile)     = n rounded up to a multiple of tile
 *   slice_rows       : P (R x C, ld) -> planes S [8][pad(R,128)][pad(C,128)] int8, scale [pad(R,128)]
 *                      (left operand of the forward product: P = A, contraction over C)
 *   slice_cols       : selected columns of P (R x C, ld) -> TRANSPOSED planes S [8][pad(ncols,tile)][pad(R,128)],
 *                      scale [pad(ncols,tile)]; colmap == NULL selects columns 0..ncols-1; tile = 128 for a left
 *                      operand (P = A for the adjoint product), 64 for a right operand (P = X or R);
 *                      scratch = ncols * 8 bytes
 *   gemm             : C[m][colmap ? colmap[n] : n] = sum_k L[m][k] R[k][n], m < Mg, n < Ng, from the planes of
 *                      L (row-scaled, pad 128) and R (column-scaled transposed, pad 64); with splits > 1, C holds
 *                      [splits][split_stride] partial products (use fb200_ozaki_splits, which also bounds the
 *                      split length so that the int32 accumulation is exact)                                */
int64_t fb200_ozaki_pad(int64_t n, int tile);
int fb200_ozaki_splits(int64_t Mg, int64_t Ng, int64_t K);
int fb200_ozaki_slice_rows(const double* P, int64_t ld, int64_t R, int64_t C, void* S, double* scale, void* stream);
int fb200_ozaki_slice_cols(const double* P, int64_t ld, int64_t R, const int* colmap, int64_t ncols, int tile, void* S,
                           double* scale, void* scratch, void* stream);
int fb200_ozaki_gemm(const void* LS, const double* lscale, int64_t Mg, const void* RS, const double* rscale, int64_t Ng,
                     int64_t K, double* C, int64_t ldc, const int* colmap, int splits, int64_t split_stride,
                     void* stream);

/* per-column vector kernels of the batched loop: arrays are (rows x B) row-major, batch fastest;
 * tau / p0 / p1 are per-column (device, B doubles), act is a per-column int mask (only active
 * columns are touched); out receives per-column sums as [k][B]:
 *   fbs_step: k = <Dx,g0>, <Dx,Dx>, |x1-xhat|^2, sum|x1|     (reference __init__.py:181-186,200,274,285)
 *   loss    : k = raw f                                        (sparse_least_squares.py:41-42)
 *   bb      : k = <Dx,Dg>, <Dg,Dg>, <g1,g1>                    (__init__.py:254-260,274)
 * ws: fb200_batched_workspace_bytes(M, N, B) bytes of device scratch.                             */
size_t fb200_batched_workspace_bytes(int64_t M, int64_t N, int64_t B);
int fb200_batched_fbs_step(const double* x0, const double* g0, const double* tau, int prox, const double* p0,
                           const double* p1, const int* act, int64_t n, int64_t B, double* xhat, double* x1,
                           double* dx, double* out, void* ws, void* stream);
int fb200_batched_loss(int loss, const double* zsrc, int nsplit, int64_t split_stride, const double* b,
                       int64_t b_ld, const int* act, int64_t m, int64_t B, double* z, double* r, double* out,
                       void* ws, void* stream);
int fb200_batched_bb(const double* gsrc, int nsplit, int64_t split_stride, const double* x0, const double* xhat,
                     const double* dx, const double* tau, const int* act, int bb, int64_t n, int64_t B,
                     double* g1, double* out, void* ws, void* stream);
int fb200_batched_select(double* dst, const double* src, const int* mask, int64_t n, int64_t B, void* stream);
/* dst[:, dcol[k]] = src[:, scol[k]], k < npairs (device index arrays): moves a column's state between its own
 * slot and the spare slots in which the batched line search evaluates several shrunken step sizes at once   */
int fb200_batched_copy_cols(double* dst, const double* src, int64_t rows, int64_t ld_dst, int64_t ld_src,
                            const int* scol, const int* dcol, int npairs, void* stream);

/* ---- K11/K12: total-variation stencils (periodic)               tv_denoising.py:26-63
 * Y is n0 x n1 x 2 (last axis interleaved), Z is n0 x n1.
 * div:  Z = sum_d roll(Y[...,d],-1,d) - Y[...,d], fused with the loss epilogue like gemv_loss.
 * grad: G[...,d] = roll(R,+1,d) - R, fused with the BB epilogue like gemvT_bb.               */
int fb200_tv_div_loss(const double* Y, int64_t n0, int64_t n1, int loss, const double* b, double* z,
                      double* r, double* scal, void* ws, void* stream);
int fb200_tv_grad_bb(const double* R, int64_t n0, int64_t n1, double* g, int bb, const double* x0,
                     const double* xhat, const double* dx, double tau, double* scal, void* ws,
                     void* stream);

/* fused TV iteration for the non-accelerated modes (K1-K3 + K11 + K5 in one pass, K11 + K8 in one):
 *   step_div_loss : xhat = x0 - tau*g0, x1 = xhat/max(|xhat|_2,1), r = div(x1) - b;
 *                   scal[S_DX_G0, S_DX_SQ, S_XMXH_SQ, S_F]   (__init__.py:181-188, tv_denoising.py:43-63,85-96)
 *   grad_bb_fused : g = grad(r); BB sums with xhat and dx recomputed from x0, g0, x1 (__init__.py:248-260) */
int fb200_tv_step_div_loss(const double* x0, const double* g0, double tau, int64_t n0, int64_t n1, int loss,
                           const double* b, double* x1, double* r, double* scal, void* ws, void* stream);
int fb200_tv_grad_bb_fused(const double* R, int64_t n0, int64_t n1, double* g, int bb, const double* x0,
                           const double* g0, const double* x1, double tau, double* scal, void* ws,
                           void* stream);

/* rank-generic forms of the same pair and of the ball projection (the reference's grad / div / prox are written for
 * N-d arrays, tv_denoising.py:26-63,89-96): X has `rank` axes `shape[0..rank)`, G / Y carry a trailing axis of size `rank`
 * (size k for the projection); one subtraction per output, the numpy summation order -- bit-identical.          */
int fb200_tv_grad_nd(const double* X, const int64_t* shape, int rank, double* G, void* stream);
int fb200_tv_div_nd(const double* Y, const int64_t* shape, int rank, double* D, void* stream);
int fb200_tv_ball_nd(const double* Y, int64_t npix, int k, double* out, void* stream);

/* whole TV iteration in one pass: step_div_loss plus the speculative g1 = grad(r) and the BB sums
 * scal[S_DX_G0, S_DX_SQ, S_XMXH_SQ, S_F, S_DX_DG, S_DG_SQ, S_G1_SQ]; reads x0, g0, b and writes x1, g1 only
 * (9U bytes, U = n0*n1*8; reference __init__.py:181-188,248-260 with tv_denoising.py:26-63,85-96)        */
int fb200_tv_iter_fused(const double* x0, const double* g0, double tau, int64_t n0, int64_t n1, int loss,
                        const double* b, double* x1, double* g1, double* scal, void* ws, void* stream);

/* ---- device-side trial decision (run-ahead without a host round trip) --------------------------------------
 * fb200_trial_decide, queued behind the kernels of a trial, repeats on the device what the host loop does with the
 * trial's sums, in the same np.float64 algebra: the non-monotone line-search test of reference __init__.py:195-201
 * (bt = backtracks already spent on this iteration), and for an accepted trial the Barzilai-Borwein step size
 * (:253-270; adaptive = 0: tau1 = tau0), the residuals (:272-281) and a built-in stop rule (stop_rule 0..3 =
 * stopping.residual / norm_residual / ratio_residual / hybrid_residual, -1 = decided by the host only).  It leaves
 * the next step size in scal[FB200_S_TAU] and sets scal[FB200_S_SKIP] when the trial was rejected or the rule fired.
 * Every entry point that takes `double tau` (fb200_fbs_step, fb200_dense_sweep, fb200_bb_reduce,
 * fb200_peer_allreduce_bb, fb200_tv_iter_fused, fb200_trial_decide) accepts tau = NaN, meaning "speculative trial":
 * read the step size from scal[FB200_S_TAU], and return at once if scal[FB200_S_SKIP] is set.  The host can
 * therefore queue the next iteration's trial before it has seen this one's sums; a trial that should not have run
 * costs a few empty launches and reports scal[FB200_S_SKIPPED] = 1.  With tau = NaN fb200_fbs_step forms the shrink
 * threshold as scal[FB200_S_TAU] * p1 (pass mu in p1).  fb200_decide_init arms the state before the first trial
 * (f0 = f_hist[0], g0_sq = |gradf(x0)|^2).  window <= FB200_FRING.  A trial queued by value (tau not NaN) re-arms the
 * device's loop state from the host's (host_it = loop index i, host_max_residual, host_g0_sq = |gradf0|^2), so a
 * last-bit disagreement between the two algebras can cost a wasted or repeated trial but never a wrong result: the
 * host decides from the sums it reads, uses scal[FB200_S_TAU_USED] as the step size of a speculative trial, and
 * repeats by value a trial that reports FB200_S_SKIPPED.  host_out (or NULL): FB200_NSCAL doubles of pinned,
 * device-visible HOST memory that receive the scalar block after the decisions -- the per-trial snapshot without a
 * memcpy in the stream (the caller waits on an event recorded behind the kernel).                                */
int fb200_decide_init(double* scal, double f0, double g0_sq, void* stream);
/* the snapshot of a trial's scalar block off the compute stream: event on `main_stream`, wait + D2H copy + event on
 * `side_stream` (src_dev = the device slot the deciding kernel filled through its host_out argument, which may be any
 * device-visible address).  Streams / events are cudaStream_t / cudaEvent_t handles owned by the caller.        */
int fb200_snapshot_copy(void* dst_host, const void* src_dev, size_t bytes, void* main_stream, void* side_stream,
                        void* ev_main, void* ev_done);
int fb200_trial_decide(double* scal, double tau, int loss, int adaptive, int backtrack, int bt, int max_backtracks,
                       int window, int stop_rule, double tolerance, int host_it, double host_max_residual,
                       double host_g0_sq, double* host_out, void* stream);

/* whole accelerated (FISTA) TV trial in one pass (reference __init__.py:181-188,220-260 with tv_denoising.py:26-63,
 * 85-96): prox point xa1 and its image za1 = div(xa1), extrapolation x1 = xa1 + c (xa1 - xa0), z1 = za1 + c (za1 - za0),
 * g1 = grad(gradf(z1)); scal[S_F] = raw f(za1) (line search, :200), scal[S_AUX3] = raw f(z1) (:245),
 * scal[S_RESTART] = <x0 - xa1, xa1 - xa0> (:231), scal[S_XMXH_SQ] with the extrapolated x1 (:274), BB sums as above.
 * c is the extrapolation weight (alpha0 - 1) / alpha1 the caller expects; if the restart test of this trial says
 * otherwise the caller repeats the call with c = 0.  15U bytes.                                                    */
int fb200_tv_fista_fused(const double* x0, const double* g0, double tau, double c, int64_t n0, int64_t n1, int loss,
                         const double* b, const double* xa0, const double* za0, double* xa1, double* za1,
                         double* x1, double* g1, double* scal, void* ws, void* stream);

/* ---- multi-GPU (A row-partitioned): all-reduce of the A^T r partials over NVLink peer memory fused
 * with the BB epilogue.  peer_ptrs[k] (host array, P entries) is the address, in THIS process, of rank k's partial:
 * n doubles of gradient partial followed by one double of raw loss partial.  g = sum_k partial_k in rank order
 * (bit-identical on every rank), scal[S_F] = sum of the loss partials at element n (with_loss >= 1; with_loss = 2 also
 * scal[S_AUX3] = sum of the second loss partials at element n + 1: the FISTA sweep), BB sums as fb200_bb_reduce.
 * The caller maps the buffers (torch symmetric memory / cudaIpc) and barriers the ranks before the call.
 * Replaces ncclAllReduce + fb200_bb_reduce for reference __init__.py:248,254-260 on the sharded map.       */
#define FB200_MAX_PEERS 16
int fb200_peer_allreduce_bb(const uint64_t* peer_ptrs, int P, int64_t n, double* g, int bb, const double* x0,
                            const double* xhat, const double* dx, double tau, int with_loss, double* scal, void* ws,
                            void* stream);

/* row-wise prox operators on a rows x cols row-major matrix iterate:
 *   mode 0: X_i * shrink(|X_i|_2, p) / (|X_i|_2 + (|X_i|_2 == 0))      prox of p * sum_i |X_i|_2   mmv.py:53-61
 *   mode 1: p * X_i / (max(|X_i|_2, p) + (|X_i|_2 == 0))               rows onto the p-ball        max_norm.py:53-59
 * out == NULL skips the prox; norms (optional, rows doubles) receives |X_i|_2                              */
int fb200_prox_rows(const double* x, int64_t rows, int64_t cols, int mode, double p, double* out, double* norms,
                    void* stream);

/* singular-value soft threshold  out = U diag(max(s - t, 0)) V  of an M x N row-major matrix: the reference's
 * proximal.project_Lnuc_ball (proximal.py:44-55, la.svd + shrink + U @ S @ V; used by
 * examples/logistic_matrix_completion.py:42-45), by one-sided Jacobi in one CTA -- no library SVD.
 * svals (optional, min(M,N) doubles) receives the singular values, unsorted; out == NULL computes them only.
 * scratch: fb200_prox_nuclear_scratch_doubles(M, N) doubles (0 = the matrix fits shared memory, pass NULL).
 * info (optional, 2 ints, device): sweeps used, converged flag.                                              */
size_t fb200_prox_nuclear_scratch_doubles(int64_t M, int64_t N);
int fb200_prox_nuclear(const double* X, int64_t M, int64_t N, int64_t ldx, double t, double* out, int64_t ldo,
                       double* svals, double* scratch, int* info, void* stream);

/* ---- small reductions used by the prologue and the generic (untagged-callable) path ---------
 * out (device) receives: dot = <a,b>; diff_nrm2sq = |a-b|^2; asum = sum |a|                   */
int fb200_dot(const double* a, const double* b, int64_t n, double* out, void* ws, void* stream);
int fb200_diff_nrm2sq(const double* a, const double* b, int64_t n, double* out, void* ws, void* stream);
int fb200_asum(const double* a, int64_t n, double* out, void* ws, void* stream);
/* out = max |a_i|  (np.abs(x).max(), reference democratic_representation.py:43; nan propagates as in numpy) */
int fb200_amax(const double* a, int64_t n, double* out, void* stream);

/* ---- row-sharded map, fused: the single-pass sweep on this rank's rows followed by ONE kernel that finishes it across
 * the ranks (csrc/dense_sweep.cu, vector_kernels.cu: peer_exchange_kernel) -- band-partial sum into this rank's slice of
 * the NVLink-mapped exchange buffer, per-chunk flags to the peers, wait for theirs, sum over ranks in rank order, BB
 * sums, and (decide_i != NULL) the decisions of fb200_trial_decide in the block that finishes last.  Replaces
 * {fb200_dense_sweep, scalar copy, barrier, fb200_peer_allreduce_bb, fb200_trial_decide} of the row-sharded iteration
 * (reference __init__.py:187-188,248,254-260 on SURVEY.md 8e's partition).
 *   peer_data[k]   address in THIS process of rank k's data slot for this call: N doubles + 2 loss partials
 *                  (two slots alternate from call to call);  peer_data[rank] is this rank's own
 *   peer_flags[k]  address of rank k's flag table: P x 128 32-bit words, zero-initialised, never written by the host
 *   epoch          call counter, identical on every rank, +1 per call (also for calls that return at once)
 *   za0 / c / za1  NULL / 0 / NULL, or the FISTA mode of fb200_dense_sweep_accel
 *   decide_i       {loss, adaptive, backtrack, bt, max_backtracks, window, stop_rule, host_it} or NULL
 *   decide_d       {tolerance, host_max_residual, host_g0_sq}
 *   host_out       as for fb200_trial_decide (used only with decide_i)                                              */
int fb200_sweep_exchange_supported(void);
int fb200_dense_sweep_exchange(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss,
                               const double* b, double* z, double* r, const double* za0, double c, double* za1,
                               const uint64_t* peer_data, const uint64_t* peer_flags, int rank, int P, uint32_t epoch,
                               double* g, int bb, const double* x0, const double* xhat, const double* dx, double tau,
                               const int* decide_i, const double* decide_d, double* host_out, double* scal, void* ws,
                               size_t ws_bytes, void* stream);

/* ---- np.random.randn on the device (csrc/legacy_rng.cu) --------------------------------------------------------
 * Replaces `x1 = np.random.randn(*x0.shape); x2 = np.random.randn(*x0.shape)` (reference fasta/__init__.py:102-103):
 * continues numpy's legacy MT19937 + polar-method Gaussian stream bit for bit from `state` and returns the state
 * numpy would be left in, so the host generator can be kept in step with np.random.set_state().
 *   state      device, 628 32-bit words: key[624], pos, has_gauss, cached gauss (a double at word 626)
 *   out        device, n doubles
 *   scratch    device, >= fb200_randn_scratch_bytes(n) bytes, 256-byte aligned
 *   state_out  device, 632 words: the end state (628 words), status (word 628: 0 ok / 1 too few accepted candidate
 *              points -- then nothing may be used; probability < 1e-11), tries used (64-bit at word 630)
 *   jump_polys device, [4][16][624] 32-bit words (fasta/mt19937_jump.npz: t^J mod phi for J = 64 * 16^level * digit
 *              MT19937 blocks, tools/make_mt_jump.py), or NULL: with it, draws beyond ~0.5M values generate the word
 *              stream from many thread blocks (jump-ahead) instead of one */
size_t fb200_randn_scratch_bytes(int64_t n);
int fb200_randn_legacy(const void* state, int64_t n, double* out, void* scratch, size_t scratch_bytes,
                       void* state_out, const void* jump_polys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTA_B200_H */

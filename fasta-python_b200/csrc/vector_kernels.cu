// N-/M-vector kernels of the FBS loop: forward step + prox + reductions (K1-K3), FISTA
// extrapolation (K9), loss evaluation (K5), Barzilai-Borwein reductions (K8), the l1-ball
// threshold search, and small reduction helpers.  All HBM-bound streaming kernels with fused
// warp-shuffle/block reductions; compiled with -fmad=false so that every elementwise expression
// rounds once per numpy operation of the reference line it replaces.
#include "common.cuh"

namespace fb200 {

// =================================================================================================
// K1-K3  xhat = x0 - tau*g0 ; x1 = prox(xhat) ; dx = x1 - x0       reference __init__.py:181-186
// =================================================================================================
template <int PROX, bool RESTART>
__global__ void __launch_bounds__(VEC_THREADS)
fbs_step_kernel(const double* __restrict__ x0, const double* __restrict__ g0, double tau, double p0,
                double p1, const double* __restrict__ xa_prev, int64_t n, double* __restrict__ xhat,
                double* __restrict__ x1, double* __restrict__ dx, double* scal, double* red,
                unsigned* counter) {
    if (PROX == FB200_PROX_L1BALL) p0 = scal[FB200_S_THETA];
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) return;
        tau = __ldcg(&scal[FB200_S_TAU]);
        if (PROX == FB200_PROX_SHRINK) p0 = tau * p1;       // proxg(x, t) = shrink(x, t * mu), mu passed in p1
    }
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a  = x0[i];
        const double gr = g0[i];
        const double h  = a - tau * gr;
        const double y  = prox_elem<PROX>(h, p0, p1);
        const double d  = y - a;
        xhat[i]         = h;
        x1[i]           = y;
        dx[i]           = d;
        s[0] += d * gr;
        s[1] += d * d;
        const double e = y - h;
        s[2] += e * e;
        s[3] += fabs(y);
        if (RESTART) s[4] += (a - y) * (y - xa_prev[i]);
    }
    double* const out[5] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ,
                            scal + FB200_S_PEN, RESTART ? scal + FB200_S_RESTART : nullptr};
    grid_sum<5>(s, red, counter, out);
}

// TV dual-ball prox works on interleaved pairs (y0, y1) / max(|y|_2, 1)   tv_denoising.py:89-96
template <bool RESTART>
__global__ void __launch_bounds__(VEC_THREADS)
fbs_step_pairs_kernel(const double2* __restrict__ x0, const double2* __restrict__ g0, double tau,
                      const double2* __restrict__ xa_prev, int64_t npairs, double2* __restrict__ xhat,
                      double2* __restrict__ x1, double2* __restrict__ dx, double* scal, double* red,
                      unsigned* counter) {
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npairs; i += stride) {
        const double2 a  = x0[i];
        const double2 gr = g0[i];
        double2 h, y, d;
        h.x = a.x - tau * gr.x;
        h.y = a.y - tau * gr.y;
        // la.norm(Y, axis=-1): sqrt(y0^2 + y1^2); then maximum(., 1); then Y / norms
        const double nrm = fmax(sqrt(h.x * h.x + h.y * h.y), 1.0);
        y.x = h.x / nrm;
        y.y = h.y / nrm;
        d.x = y.x - a.x;
        d.y = y.y - a.y;
        xhat[i] = h;
        x1[i]   = y;
        dx[i]   = d;
        s[0] += d.x * gr.x;
        s[0] += d.y * gr.y;
        s[1] += d.x * d.x;
        s[1] += d.y * d.y;
        const double e0 = y.x - h.x, e1 = y.y - h.y;
        s[2] += e0 * e0;
        s[2] += e1 * e1;
        if (RESTART) {
            const double2 q = xa_prev[i];
            s[4] += (a.x - y.x) * (y.x - q.x);
            s[4] += (a.y - y.y) * (y.y - q.y);
        }
    }
    double* const out[5] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ,
                            scal + FB200_S_PEN, RESTART ? scal + FB200_S_RESTART : nullptr};
    grid_sum<5>(s, red, counter, out);
}

template <int PROX>
static int launch_fbs(const double* x0, const double* g0, double tau, double p0, double p1,
                      const double* xa_prev, int64_t n, double* xhat, double* x1, double* dx,
                      double* scal, Workspace& w, cudaStream_t st) {
    const int grid = vec_grid(n);
    if (xa_prev)
        fbs_step_kernel<PROX, true><<<grid, VEC_THREADS, 0, st>>>(x0, g0, tau, p0, p1, xa_prev, n, xhat,
                                                                  x1, dx, scal, w.red, w.counter);
    else
        fbs_step_kernel<PROX, false><<<grid, VEC_THREADS, 0, st>>>(x0, g0, tau, p0, p1, nullptr, n, xhat,
                                                                   x1, dx, scal, w.red, w.counter);
    return check_launch("fbs_step");
}

// generic-path pieces: forward step alone, reductions alone, prox alone
__global__ void __launch_bounds__(VEC_THREADS)
forward_step_kernel(const double* __restrict__ x0, const double* __restrict__ g0, double tau, int64_t n,
                    double* __restrict__ xhat) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        xhat[i] = x0[i] - tau * g0[i];
}

template <bool RESTART>
__global__ void __launch_bounds__(VEC_THREADS)
step_reduce_kernel(const double* __restrict__ x0, const double* __restrict__ x1, const double* __restrict__ xhat,
                   const double* __restrict__ g0, const double* __restrict__ xa_prev, int64_t n,
                   double* __restrict__ dx, double* scal, double* red, unsigned* counter) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = x0[i], y = x1[i];
        const double d = y - a;
        dx[i]          = d;
        s[0] += d * g0[i];
        s[1] += d * d;
        const double e = y - xhat[i];
        s[2] += e * e;
        if (RESTART) s[3] += (a - y) * (y - xa_prev[i]);
    }
    double* const out[4] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ,
                            RESTART ? scal + FB200_S_RESTART : nullptr};
    grid_sum<4>(s, red, counter, out);
}

template <int PROX>
__global__ void __launch_bounds__(VEC_THREADS)
prox_apply_kernel(const double* __restrict__ x, double p0, double p1, int64_t n, double* __restrict__ out,
                  const double* __restrict__ scal) {
    if (PROX == FB200_PROX_L1BALL) p0 = scal[FB200_S_THETA];
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    if (PROX == FB200_PROX_TV_BALL) {
        const double2* xp = reinterpret_cast<const double2*>(x);
        double2* op       = reinterpret_cast<double2*>(out);
        for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n / 2; i += stride) {
            const double2 h  = xp[i];
            const double nrm = fmax(sqrt(h.x * h.x + h.y * h.y), 1.0);
            op[i]            = make_double2(h.x / nrm, h.y / nrm);
        }
    } else {
        for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
            out[i] = prox_elem<PROX>(x[i], p0, p1);
    }
}

// =================================================================================================
// l1-ball threshold (Michelot's fixed point; exact active set in finitely many passes)
//   theta solves sum_i max(|v_i| - theta, 0) = radius; equals the reference's
//   alpha = max_k (cumsum_k(sorted |v|) - t)/k                     proximal.py:22-26
// One CTA; each pass is a fixed-order block reduction, so the result is bit-reproducible.
// =================================================================================================
constexpr int L1B_THREADS = 1024;

__global__ void __launch_bounds__(L1B_THREADS)
l1ball_threshold_kernel(const double* __restrict__ v, int64_t n, double radius, double* scal) {
    __shared__ double sm[2 * 32];
    __shared__ double sh_theta;
    __shared__ double sh_count;
    double theta = -1.0;    // first pass: every element is active (|v| > -1)
    double prev_count = -1.0;
    // Michelot's fixed point: every pass drops at least one element from the active set or stops, so n + 1 passes bound it
    // (typically ~10); no artificial cap -- a capped run would silently return a wrong threshold on adversarial input
    for (int64_t pass = 0; pass <= n + 1; ++pass) {
        double s[2] = {0.0, 0.0};
        for (int64_t i = threadIdx.x; i < n; i += L1B_THREADS) {
            const double m = fabs(v[i]);
            if (m > theta) {
                s[0] += m;
                s[1] += 1.0;
            }
        }
        block_sum<2>(s, sm);
        if (threadIdx.x == 0) {
            sh_count = s[1];
            if (pass == 0 && s[0] <= radius) {
                sh_theta = 0.0;      // already inside the ball: projection is the identity
                sh_count = -2.0;     // sentinel: stop
            } else {
                sh_theta = (s[1] > 0.0) ? (s[0] - radius) / s[1] : theta;
            }
        }
        __syncthreads();
        const double cnt = sh_count;
        theta            = sh_theta;
        if (cnt == -2.0 || cnt == prev_count || cnt <= 0.0) break;
        prev_count = cnt;
        __syncthreads();
    }
    if (threadIdx.x == 0) scal[FB200_S_THETA] = (theta > 0.0) ? theta : 0.0;
}

// =================================================================================================
// K9  FISTA extrapolation of x (n) and z (m) + loss at the extrapolated z  __init__.py:242-245
// =================================================================================================
template <int LOSS, bool PEN>
__global__ void __launch_bounds__(VEC_THREADS)
accel_step_kernel(double c, const double* __restrict__ xa1, const double* __restrict__ xa0,
                  const double* __restrict__ xhat, int64_t n, double* __restrict__ x1,
                  const double* __restrict__ za1, const double* __restrict__ za0,
                  const double* __restrict__ b, int64_t m, double* __restrict__ z1,
                  double* __restrict__ r, double* scal, double* red, unsigned* counter) {
    double s[3] = {0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    const int64_t t0     = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n; i += stride) {
        const double p = xa1[i];
        const double y = p + c * (p - xa0[i]);          // x1 + (alpha0-1)/alpha1 * (xa1 - xa0)
        x1[i]          = y;
        const double e = y - xhat[i];
        s[1] += e * e;
        if (PEN) s[2] += fabs(y);
    }
    for (int64_t i = t0; i < m; i += stride) {
        const double p = za1[i];
        const double z = p + c * (p - za0[i]);
        z1[i]          = z;
        double ri, fi;
        loss_elem<LOSS>(z, b ? b[i] : 0.0, ri, fi);
        if (LOSS != FB200_LOSS_NONE) r[i] = ri;
        s[0] += fi;
    }
    double* const out[3] = {scal + FB200_S_F, scal + FB200_S_XMXH_SQ, scal + FB200_S_PEN};
    grid_sum<3>(s, red, counter, out);
}

// =================================================================================================
// K5  z (optionally the fixed-order sum of S split partials) -> r = gradf(z), raw f(z)
// =================================================================================================
template <int LOSS>
__global__ void __launch_bounds__(VEC_THREADS)
loss_kernel(const double* __restrict__ zsrc, int nsplit, int64_t ld, const double* __restrict__ b,
            int64_t m, double* __restrict__ z, double* __restrict__ r, double* scal, double* red,
            unsigned* counter) {
    double s[1] = {0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += stride) {
        double zi;
        if (nsplit > 0) {
            zi = __ldcg(&zsrc[i]);
            for (int k = 1; k < nsplit; ++k) zi += __ldcg(&zsrc[int64_t(k) * ld + i]);
            z[i] = zi;
        } else {
            zi = zsrc[i];
        }
        if (LOSS != FB200_LOSS_NONE) {
            double ri, fi;
            loss_elem<LOSS>(zi, b[i], ri, fi);
            r[i] = ri;
            s[0] += fi;
        }
    }
    if (LOSS != FB200_LOSS_NONE) {
        double* const out[1] = {scal + FB200_S_F};
        grid_sum<1>(s, red, counter, out);
    }
}

int launch_loss(int loss, const double* zsrc, int nsplit, int64_t ld, const double* b, int64_t m,
                double* z, double* r, double* scal, Workspace& w, cudaStream_t st) {
    const int grid = vec_grid(m);
    switch (loss) {
        case FB200_LOSS_NONE:
            if (nsplit == 0) return 0;
            loss_kernel<FB200_LOSS_NONE><<<grid, VEC_THREADS, 0, st>>>(zsrc, nsplit, ld, b, m, z, r, scal, w.red, w.counter);
            break;
        case FB200_LOSS_LEAST_SQUARES:
            loss_kernel<FB200_LOSS_LEAST_SQUARES><<<grid, VEC_THREADS, 0, st>>>(zsrc, nsplit, ld, b, m, z, r, scal, w.red, w.counter);
            break;
        case FB200_LOSS_LOGISTIC:
            loss_kernel<FB200_LOSS_LOGISTIC><<<grid, VEC_THREADS, 0, st>>>(zsrc, nsplit, ld, b, m, z, r, scal, w.red, w.counter);
            break;
        default:
            set_error("unknown loss tag %d", loss);
            return 1;
    }
    return check_launch("loss_kernel");
}

// =================================================================================================
// K8  g1 (optionally summed from S split partials) ; dg = g1 + (xhat - x0)/tau ; reductions
// =================================================================================================
template <int BB>
__global__ void __launch_bounds__(VEC_THREADS)
bb_kernel(const double* __restrict__ gsrc, int nsplit, int64_t ld, int64_t n, double* __restrict__ g,
          const double* __restrict__ x0, const double* __restrict__ xhat, const double* __restrict__ dx,
          double tau, double* scal, double* red, unsigned* counter) {
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) return;
        tau = __ldcg(&scal[FB200_S_TAU]);
    }
    double s[3] = {0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double gi;
        if (nsplit > 0) {
            gi = __ldcg(&gsrc[i]);
            for (int k = 1; k < nsplit; ++k) gi += __ldcg(&gsrc[int64_t(k) * ld + i]);
            g[i] = gi;
        } else {
            gi = gsrc[i];
        }
        if (BB >= 1) s[2] += gi * gi;
        if (BB >= 2) {
            const double dg = gi + (xhat[i] - x0[i]) / tau;     // __init__.py:254
            s[0] += dx[i] * dg;                                  // :255
            s[1] += dg * dg;                                     // :260
        }
    }
    if (BB >= 1) {
        double* const out[3] = {BB >= 2 ? scal + FB200_S_DX_DG : nullptr,
                                BB >= 2 ? scal + FB200_S_DG_SQ : nullptr, scal + FB200_S_G1_SQ};
        grid_sum<3>(s, red, counter, out);
    }
}

int launch_bb(int bb, const double* gsrc, int nsplit, int64_t ld, int64_t n, double* g, const double* x0,
              const double* xhat, const double* dx, double tau, double* scal, Workspace& w,
              cudaStream_t st) {
    const int grid = vec_grid(n);
    switch (bb) {
        case 0:
            if (nsplit == 0) return 0;
            bb_kernel<0><<<grid, VEC_THREADS, 0, st>>>(gsrc, nsplit, ld, n, g, x0, xhat, dx, tau, scal, w.red, w.counter);
            break;
        case 1:
            bb_kernel<1><<<grid, VEC_THREADS, 0, st>>>(gsrc, nsplit, ld, n, g, x0, xhat, dx, tau, scal, w.red, w.counter);
            break;
        case 2:
            bb_kernel<2><<<grid, VEC_THREADS, 0, st>>>(gsrc, nsplit, ld, n, g, x0, xhat, dx, tau, scal, w.red, w.counter);
            break;
        default:
            set_error("unknown bb mode %d", bb);
            return 1;
    }
    return check_launch("bb_kernel");
}

// ---- multi-GPU: all-reduce of the row-sharded A^T r partials over peer memory, fused with the BB epilogue ----
// Every rank's partial gradient (n doubles) and raw loss partial (1 double at index n) sit in a buffer that is
// mapped into every other rank's address space (NVLink peer memory; the mapping and the barrier before this
// kernel are torch symmetric-memory plumbing).  Each rank reads all P partials directly from its peers and adds
// them in rank order -- so all ranks hold the bit-identical g -- and forms the Barzilai-Borwein sums in the same
// pass (reference __init__.py:248 + :254-260,274 on the row-partitioned map of SURVEY.md 8e).  One kernel
// replaces {copy, ncclAllReduce, copy, bb_kernel}.
struct PeerParts { const double* p[FB200_MAX_PEERS]; };

template <int BB>
__global__ void __launch_bounds__(VEC_THREADS)
peer_allreduce_bb_kernel(PeerParts parts, int P, int64_t n, double* __restrict__ g, const double* __restrict__ x0,
                         const double* __restrict__ xhat, const double* __restrict__ dx, double tau, int with_loss,
                         double* scal, double* red, unsigned* counter) {
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) return;
        tau = __ldcg(&scal[FB200_S_TAU]);
    }
    double s[3] = {0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double gi = __ldcg(&parts.p[0][i]);
        for (int k = 1; k < P; ++k) gi += __ldcg(&parts.p[k][i]);
        g[i] = gi;
        if (BB >= 1) s[2] += gi * gi;
        if (BB >= 2) {
            const double dg = gi + (xhat[i] - x0[i]) / tau;     // __init__.py:254
            s[0] += dx[i] * dg;                                  // :255
            s[1] += dg * dg;                                     // :260
        }
    }
    if (with_loss && blockIdx.x == 0 && threadIdx.x == 0) {
        double f = __ldcg(&parts.p[0][n]);
        for (int k = 1; k < P; ++k) f += __ldcg(&parts.p[k][n]);
        scal[FB200_S_F] = f;
        if (with_loss >= 2) {                               // FISTA sweep: second loss partial (extrapolated point)
            double f2 = __ldcg(&parts.p[0][n + 1]);
            for (int k = 1; k < P; ++k) f2 += __ldcg(&parts.p[k][n + 1]);
            scal[FB200_S_AUX3] = f2;
        }
    }
    if (BB >= 1) {
        double* const out[3] = {BB >= 2 ? scal + FB200_S_DX_DG : nullptr,
                                BB >= 2 ? scal + FB200_S_DG_SQ : nullptr, scal + FB200_S_G1_SQ};
        grid_sum<3>(s, red, counter, out);
    }
}

// =================================================================================================
// The host loop's decisions on the device (see fb200_trial_decide in the header): line search (:195-201), step size
// (:253-270), residuals (:272-281), stop rule (:308).  la.norm(.)**2 is sqrt(.)**2; Python max(a, b) is (b > a ? b : a).
// One thread, queued behind the trial's kernels.
// =================================================================================================
__device__ __forceinline__ double dd_sq(double v) { const double n = sqrt(v); return n * n; }
__device__ __forceinline__ double dd_pymax(double a, double b) { return (b > a) ? b : a; }

__global__ void decide_init_kernel(double* scal, double f0, double g0_sq) {
    scal[FB200_S_SKIP] = 0.0;
    scal[FB200_S_SKIPPED] = 0.0;
    scal[FB200_S_IT] = 0.0;
    scal[FB200_S_MAXRES] = -INFINITY;
    scal[FB200_S_G0SQ] = g0_sq;
    scal[FB200_S_FRING] = f0;
}

struct DecideArgs {
    int loss, adaptive, backtrack, bt, max_backtracks, window, stop_rule, host_it;
    double tolerance, host_max_residual, host_g0_sq;
};

__device__ void decide_core(double* scal, double tau, const DecideArgs& d) {
    const bool speculative = isnan(tau);
    if (speculative && scal[FB200_S_SKIP] != 0.0) { scal[FB200_S_SKIPPED] = 1.0; return; }
    scal[FB200_S_SKIPPED] = 0.0;
    if (!speculative) {                                     // a trial queued by value carries the host's loop state
        scal[FB200_S_IT] = double(d.host_it);
        scal[FB200_S_MAXRES] = d.host_max_residual;
        scal[FB200_S_G0SQ] = d.host_g0_sq;
    }
    const double tau0 = speculative ? scal[FB200_S_TAU] : tau;
    scal[FB200_S_TAU_USED] = tau0;
    const double raw = scal[FB200_S_F];
    const double f1 = (d.loss == FB200_LOSS_LEAST_SQUARES) ? .5 * dd_sq(raw) : raw;
    const int it = int(scal[FB200_S_IT]);
    const double dx_sq = scal[FB200_S_DX_SQ];
    if (d.backtrack && d.bt < d.max_backtracks) {
        double fmax_w = -INFINITY;
        bool has_nan = false;
        for (int k = (it - d.window + 1 > 0 ? it - d.window + 1 : 0); k <= it; ++k) {
            const double v = scal[FB200_S_FRING + k % FB200_FRING];
            if (isnan(v)) has_nan = true;
            fmax_w = fmax(fmax_w, v);
        }
        if (has_nan) fmax_w = NAN;                                   // np.max propagates nan
        if (f1 - (fmax_w + scal[FB200_S_DX_G0] + dd_sq(dx_sq) / (2 * tau0)) > 1E-12) {
            scal[FB200_S_SKIP] = 1.0;                                // rejected: the host backtracks (by-value trials)
            return;
        }
    }
    const double dx_norm = sqrt(dx_sq);
    double tau1 = tau0;
    if (d.adaptive) {
        const double dotprod = scal[FB200_S_DX_DG];
        const double tau_s = (dx_norm * dx_norm) / dotprod;
        const double q = dotprod / dd_sq(scal[FB200_S_DG_SQ]);
        const double tau_m = (0.0 > q) ? 0.0 : q;                    // Python max(q, 0): a nan q stays
        if (2 * tau_m > tau_s) tau1 = tau_m;
        else tau1 = tau_s - .5 * tau_m;
        if (tau1 <= 0 || isinf(tau1) || isnan(tau1)) tau1 = tau0 * 1.5;
    }
    const double resid = dx_norm / tau0;
    const double normalizer = dd_pymax(sqrt(scal[FB200_S_G0SQ]), sqrt(scal[FB200_S_XMXH_SQ]) / tau0) + 1E-12;
    const double nresid = resid / normalizer;
    const double max_residual = dd_pymax(scal[FB200_S_MAXRES], resid);
    bool stop = false;
    switch (d.stop_rule) {
        case 0: stop = resid < d.tolerance; break;
        case 1: stop = nresid < d.tolerance; break;
        case 2: stop = resid / max_residual < d.tolerance; break;
        case 3: stop = (resid / max_residual < d.tolerance) || (nresid < d.tolerance); break;
        default: break;
    }
    scal[FB200_S_FRING + (it + 1) % FB200_FRING] = f1;
    scal[FB200_S_IT] = double(it + 1);
    scal[FB200_S_MAXRES] = max_residual;
    scal[FB200_S_G0SQ] = scal[FB200_S_G1_SQ];
    scal[FB200_S_TAU] = tau1;
    scal[FB200_S_SKIP] = stop ? 1.0 : 0.0;
}

// host_out (optional): pinned, device-visible host memory that receives the scalar block once the decisions are taken --
// the per-trial snapshot without a memcpy node in the stream (the host waits on an event recorded behind the kernel)
__global__ void __launch_bounds__(FB200_NSCAL)
trial_decide_kernel(double* scal, double tau, DecideArgs d, double* __restrict__ host_out) {
    if (threadIdx.x == 0) decide_core(scal, tau, d);
    if (!host_out) return;
    __syncthreads();
    host_out[threadIdx.x] = __ldcg(&scal[threadIdx.x]);
}

// =================================================================================================
// Row-sharded map, ONE kernel for what follows the local sweep (reference __init__.py:248,254-260,195-201,253-281 on the
// row-partitioned map of SURVEY.md 8e): sum of this rank's band partials -> this rank's slice of the exchange buffer
// (NVLink-mapped on every peer) -> per-chunk flags to the peers -> wait for the peers' flags of the same chunk -> read
// their partials over NVLink, add in rank order (bit-identical g on every rank) -> Barzilai-Borwein sums -> and in the
// block that finishes last the loop's decisions (fb200_trial_decide).  Block c owns chunk c of the N-vector from its
// first instruction to its last, so the exchange of chunk c overlaps the band sums of the chunks behind it and nobody
// waits for the whole vector: there is no barrier between "produce" and "consume", only one flag per (rank, chunk).
//   flags[src][chunk] (32-bit, in every rank's mapped buffer) = epoch of the last call whose chunk `chunk` rank `src`
//   finished writing; epochs grow by one per call on every rank, a waiter accepts any epoch >= its own.
//   Two data slots alternate: a rank rewrites a slot two calls later, which it can only reach after every peer has
//   signalled the call in between, i.e. finished reading (same argument as the two-buffer barrier scheme before).
// All blocks of a rank must be resident (a block spins on its peers' flags): the grid is PEER_CHUNKS <= #SMs blocks.
// =================================================================================================
constexpr int PEER_CHUNKS = 128;

struct PeerPtrs {
    double*   data[FB200_MAX_PEERS];     // this call's slot of rank k's exchange buffer: n doubles + 2 loss partials
    uint32_t* flags[FB200_MAX_PEERS];    // rank k's flag table [P][PEER_CHUNKS]
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int BB>
__global__ void __launch_bounds__(VEC_THREADS)
peer_exchange_kernel(PeerPtrs pp, int rank, int P, uint32_t epoch, int64_t n, const double* __restrict__ gsrc, int nsplit,
                     int64_t ld, const double* __restrict__ fpart, const double* __restrict__ fpart2, int with_loss,
                     double* __restrict__ g, const double* __restrict__ x0, const double* __restrict__ xhat,
                     const double* __restrict__ dx, double tau, int decide, DecideArgs dargs, double* scal, double* red,
                     unsigned* counter, double* __restrict__ host_out) {
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) {
            if (decide && blockIdx.x == 0) {
                if (threadIdx.x == 0) decide_core(scal, tau, dargs);    // reports S_SKIPPED
                if (host_out) {
                    __syncthreads();
                    if (threadIdx.x < FB200_NSCAL) host_out[threadIdx.x] = __ldcg(&scal[threadIdx.x]);
                }
            }
            return;
        }
    }
    const double tau_v = isnan(tau) ? __ldcg(&scal[FB200_S_TAU]) : tau;
    const int c = blockIdx.x;
    const int64_t per = ((n + PEER_CHUNKS - 1) / PEER_CHUNKS + 1) / 2 * 2;      // even: chunks start 16-byte aligned
    const int64_t lo = int64_t(c) * per, hi = (lo + per < n) ? lo + per : n;    // n is even (the sweep requires it)
    double* mine = pp.data[rank];
    // 1. this rank's partial of the chunk: the band partials of the local sweep in index order (loads batched: they
    //    are independent, only the additions are ordered)
    for (int64_t i = lo + 2 * threadIdx.x; i < hi; i += 2 * blockDim.x) {
        double2 acc = __ldcg(reinterpret_cast<const double2*>(gsrc + i));
        for (int k0 = 1; k0 < nsplit; k0 += 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < nsplit) v[u] = __ldcg(reinterpret_cast<const double2*>(gsrc + int64_t(k0 + u) * ld + i));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < nsplit) { acc.x += v[u].x; acc.y += v[u].y; }
        }
        *reinterpret_cast<double2*>(mine + i) = acc;
    }
    if (c == 0 && threadIdx.x == 0 && with_loss) {
        double f = 0.0;
        for (int k = 0; k < nsplit; ++k) f += __ldcg(&fpart[k]);
        mine[n] = f;
        if (with_loss >= 2) {
            double f2 = 0.0;
            for (int k = 0; k < nsplit; ++k) f2 += __ldcg(&fpart2[k]);
            mine[n + 1] = f2;
        }
    }
    __syncthreads();
    // 2. tell every peer that chunk c of this rank is complete, 3. wait for theirs
    if (threadIdx.x < P && threadIdx.x != rank) {
        __threadfence_system();
        st_release_sys_u32(pp.flags[threadIdx.x] + rank * PEER_CHUNKS + c, epoch);
        const uint32_t* f = pp.flags[rank] + threadIdx.x * PEER_CHUNKS + c;
        while (int32_t(ld_acquire_sys_u32(f) - epoch) < 0) {}
    }
    __syncthreads();
    // 4. g = sum over ranks in rank order, Barzilai-Borwein sums.  The P remote loads of an element pair are issued
    //    together (one NVLink round trip per pair, not P): L1 is bypassed (.cg), the acquire above orders them.
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t i = lo + 2 * threadIdx.x; i < hi; i += 2 * blockDim.x) {
        double2 v[FB200_MAX_PEERS];
#pragma unroll
        for (int k = 0; k < FB200_MAX_PEERS; ++k)
            if (k < P) v[k] = __ldcg(reinterpret_cast<const double2*>(pp.data[k] + i));
        double2 gi = v[0];
#pragma unroll
        for (int k = 1; k < FB200_MAX_PEERS; ++k)
            if (k < P) { gi.x += v[k].x; gi.y += v[k].y; }
        *reinterpret_cast<double2*>(g + i) = gi;
        if (BB >= 1) { s[2] += gi.x * gi.x; s[2] += gi.y * gi.y; }
        if (BB >= 2) {
            const double2 xh = *reinterpret_cast<const double2*>(xhat + i), xo = *reinterpret_cast<const double2*>(x0 + i);
            const double2 dd = *reinterpret_cast<const double2*>(dx + i);
            const double dg0 = gi.x + (xh.x - xo.x) / tau_v;     // __init__.py:254
            const double dg1 = gi.y + (xh.y - xo.y) / tau_v;
            s[0] += dd.x * dg0; s[0] += dd.y * dg1;              // :255
            s[1] += dg0 * dg0;  s[1] += dg1 * dg1;               // :260
        }
    }
    if (with_loss && c == 0 && threadIdx.x == 0) {
        double f = __ldcg(pp.data[0] + n);
        for (int k = 1; k < P; ++k) f += __ldcg(pp.data[k] + n);
        scal[FB200_S_F] = f;
        if (with_loss >= 2) {                               // FISTA sweep: second loss partial (extrapolated point)
            double f2 = __ldcg(pp.data[0] + n + 1);
            for (int k = 1; k < P; ++k) f2 += __ldcg(pp.data[k] + n + 1);
            scal[FB200_S_AUX3] = f2;
        }
    }
    double* const out[3] = {BB >= 2 ? scal + FB200_S_DX_DG : nullptr, BB >= 2 ? scal + FB200_S_DG_SQ : nullptr,
                            BB >= 1 ? scal + FB200_S_G1_SQ : nullptr};
    const bool last = grid_sum_last<3>(s, red, counter, out);
    if (last && decide) {
        if (threadIdx.x == 0) {
            __threadfence();                                // block 0's S_F precedes its ticket; ours came after
            decide_core(scal, tau, dargs);
        }
        if (host_out) {                                     // the snapshot, straight into pinned host memory
            __syncthreads();
            if (threadIdx.x < FB200_NSCAL) host_out[threadIdx.x] = __ldcg(&scal[threadIdx.x]);
        }
    }
}

int launch_peer_exchange(const uint64_t* peer_data, const uint64_t* peer_flags, int rank, int P, uint32_t epoch, int64_t n,
                         const double* gsrc, int nsplit, int64_t ld, const double* fpart, const double* fpart2, int with_loss,
                         double* g, int bb, const double* x0, const double* xhat, const double* dx, double tau,
                         const int* decide_i, const double* decide_d, double* host_out, double* scal, Workspace& w, cudaStream_t st) {
    PeerPtrs pp;
    for (int k = 0; k < FB200_MAX_PEERS; ++k) {
        pp.data[k] = reinterpret_cast<double*>(peer_data[k < P ? k : 0]);
        pp.flags[k] = reinterpret_cast<uint32_t*>(peer_flags[k < P ? k : 0]);
    }
    DecideArgs d{};
    const int decide = decide_i ? 1 : 0;
    if (decide) {
        d.loss = decide_i[0]; d.adaptive = decide_i[1]; d.backtrack = decide_i[2]; d.bt = decide_i[3]; d.max_backtracks = decide_i[4];
        d.window = decide_i[5]; d.stop_rule = decide_i[6]; d.host_it = decide_i[7];
        d.tolerance = decide_d[0]; d.host_max_residual = decide_d[1]; d.host_g0_sq = decide_d[2];
        if (d.window > FB200_FRING) { set_error("peer_exchange: window %d exceeds the device ring (%d)", d.window, FB200_FRING); return 1; }
    }
    if (sm_count() < PEER_CHUNKS) { set_error("peer_exchange: needs %d resident blocks", PEER_CHUNKS); return 1; }
#define FB200_PEERX(BBV) peer_exchange_kernel<BBV><<<PEER_CHUNKS, VEC_THREADS, 0, st>>>(pp, rank, P, epoch, n, gsrc, nsplit, ld, fpart, fpart2, with_loss, g, x0, xhat, dx, tau, decide, d, scal, w.red, w.counter, decide ? host_out : nullptr)
    switch (bb) {
        case 0: FB200_PEERX(0); break;
        case 1: FB200_PEERX(1); break;
        case 2: FB200_PEERX(2); break;
        default: set_error("unknown bb mode %d", bb); return 1;
    }
#undef FB200_PEERX
    return check_launch("peer_exchange");
}

// =================================================================================================
// small reductions
// =================================================================================================
template <int OP>   // 0: <a,b>   1: |a-b|^2   2: sum |a|
__global__ void __launch_bounds__(VEC_THREADS)
reduce_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* out,
              double* red, unsigned* counter) {
    double s[1] = {0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (OP == 0) {
            s[0] += a[i] * b[i];
        } else if (OP == 1) {
            const double d = a[i] - b[i];
            s[0] += d * d;
        } else {
            s[0] += fabs(a[i]);
        }
    }
    double* const o[1] = {out};
    grid_sum<1>(s, red, counter, o);
}

// max |a_i| (np.abs(x).max(), reference democratic_representation.py:43): the bit patterns of non-negative doubles order
// like unsigned integers (a nan sorts above inf, so it propagates as in numpy); max is order-independent: deterministic
__global__ void amax_zero_kernel(double* out) { *out = 0.0; }
__global__ void __launch_bounds__(VEC_THREADS) amax_kernel(const double* __restrict__ a, int64_t n, double* out) {
    unsigned long long m = 0ull;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long v = (unsigned long long)__double_as_longlong(fabs(a[i]));
        m = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned long long*>(out), m);
}

// ---- row-wise prox operators of the matrix-iterate examples ------------------------------------------
// X is rows x cols row-major; one warp per row: nrm = sqrt(sum_j X[i][j]^2) (lane-strided partial sums,
// butterfly combine), then
//   mode 0 (mmv.py:53-61, prox of t*sum_i |X_i|_2):  X_i * (shrink(nrm, p) / (nrm + (nrm == 0)))
//   mode 1 (max_norm.py:53-59, rows onto the p-ball): p * X_i / (max(nrm, p) + (nrm == 0))
// norms (optional) receives nrm per row (for g = mu * sum_i |X_i|_2).
__global__ void __launch_bounds__(256)
prox_rows_kernel(const double* __restrict__ x, int64_t rows, int cols, int mode, double p, double* __restrict__ out,
                 double* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double* xr = x + row * cols;
    double s = 0.0;
    for (int j = lane; j < cols; j += 32) s += xr[j] * xr[j];
    s = warp_sum(s);
    const double nrm = sqrt(s);
    if (norms && lane == 0) norms[row] = nrm;
    if (!out) return;
    const double zero = (nrm == 0.0) ? 1.0 : 0.0;
    if (mode == 0) {
        const double scale = (sign_np(nrm) * fmax(fabs(nrm) - p, 0.0)) / (nrm + zero);
        for (int j = lane; j < cols; j += 32) out[row * cols + j] = xr[j] * scale;
    } else {
        const double scale = fmax(nrm, p) + zero;
        for (int j = lane; j < cols; j += 32) out[row * cols + j] = (p * xr[j]) / scale;
    }
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_fbs_step(const double* x0, const double* g0, double tau, int prox, double p0, double p1,
                              const double* xa_prev, int64_t n, double* xhat, double* x1, double* dx,
                              double* scal, void* ws, void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) { set_error("fbs_step: n must be positive"); return 1; }
    switch (prox) {
        case FB200_PROX_IDENTITY: return launch_fbs<FB200_PROX_IDENTITY>(x0, g0, tau, p0, p1, xa_prev, n, xhat, x1, dx, scal, w, st);
        case FB200_PROX_SHRINK:   return launch_fbs<FB200_PROX_SHRINK>(x0, g0, tau, p0, p1, xa_prev, n, xhat, x1, dx, scal, w, st);
        case FB200_PROX_NONNEG:   return launch_fbs<FB200_PROX_NONNEG>(x0, g0, tau, p0, p1, xa_prev, n, xhat, x1, dx, scal, w, st);
        case FB200_PROX_BOX:      return launch_fbs<FB200_PROX_BOX>(x0, g0, tau, p0, p1, xa_prev, n, xhat, x1, dx, scal, w, st);
        case FB200_PROX_L1BALL:   return launch_fbs<FB200_PROX_L1BALL>(x0, g0, tau, p0, p1, xa_prev, n, xhat, x1, dx, scal, w, st);
        case FB200_PROX_TV_BALL: {
            if (n % 2) { set_error("fbs_step: TV ball prox needs an even element count"); return 1; }
            const int grid = vec_grid(n / 2, 1);
            if (xa_prev)
                fbs_step_pairs_kernel<true><<<grid, VEC_THREADS, 0, st>>>(
                    (const double2*)x0, (const double2*)g0, tau, (const double2*)xa_prev, n / 2, (double2*)xhat,
                    (double2*)x1, (double2*)dx, scal, w.red, w.counter);
            else
                fbs_step_pairs_kernel<false><<<grid, VEC_THREADS, 0, st>>>(
                    (const double2*)x0, (const double2*)g0, tau, nullptr, n / 2, (double2*)xhat, (double2*)x1,
                    (double2*)dx, scal, w.red, w.counter);
            return check_launch("fbs_step_pairs");
        }
        default: set_error("unknown prox tag %d", prox); return 1;
    }
}

extern "C" int fb200_forward_step(const double* x0, const double* g0, double tau, int64_t n, double* xhat,
                                  void* stream) {
    forward_step_kernel<<<vec_grid(n), VEC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x0, g0, tau, n, xhat);
    return check_launch("forward_step");
}

extern "C" int fb200_step_reduce(const double* x0, const double* x1, const double* xhat, const double* g0,
                                 const double* xa_prev, int64_t n, double* dx, double* scal, void* ws,
                                 void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (xa_prev)
        step_reduce_kernel<true><<<vec_grid(n), VEC_THREADS, 0, st>>>(x0, x1, xhat, g0, xa_prev, n, dx, scal, w.red, w.counter);
    else
        step_reduce_kernel<false><<<vec_grid(n), VEC_THREADS, 0, st>>>(x0, x1, xhat, g0, nullptr, n, dx, scal, w.red, w.counter);
    return check_launch("step_reduce");
}

extern "C" int fb200_peer_allreduce_bb(const uint64_t* peer_ptrs, int P, int64_t n, double* g, int bb, const double* x0,
                                       const double* xhat, const double* dx, double tau, int with_loss, double* scal,
                                       void* ws, void* stream) {
    if (P < 1 || P > FB200_MAX_PEERS || n < 1) { set_error("peer_allreduce_bb: bad arguments"); return 1; }
    Workspace w(ws);
    PeerParts parts;
    for (int k = 0; k < FB200_MAX_PEERS; ++k) parts.p[k] = reinterpret_cast<const double*>(k < P ? peer_ptrs[k] : peer_ptrs[0]);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = vec_grid(n);
    switch (bb) {
        case 0: peer_allreduce_bb_kernel<0><<<grid, VEC_THREADS, 0, st>>>(parts, P, n, g, x0, xhat, dx, tau, with_loss, scal, w.red, w.counter); break;
        case 1: peer_allreduce_bb_kernel<1><<<grid, VEC_THREADS, 0, st>>>(parts, P, n, g, x0, xhat, dx, tau, with_loss, scal, w.red, w.counter); break;
        case 2: peer_allreduce_bb_kernel<2><<<grid, VEC_THREADS, 0, st>>>(parts, P, n, g, x0, xhat, dx, tau, with_loss, scal, w.red, w.counter); break;
        default: set_error("unknown bb mode %d", bb); return 1;
    }
    return check_launch("peer_allreduce_bb");
}

extern "C" int fb200_prox_rows(const double* x, int64_t rows, int64_t cols, int mode, double p, double* out, double* norms,
                               void* stream) {
    if (rows < 1 || cols < 1 || cols > INT32_MAX || (mode != 0 && mode != 1)) { set_error("prox_rows: bad arguments"); return 1; }
    const int64_t grid = (rows + 7) / 8;
    if (grid > INT32_MAX) { set_error("prox_rows: too many rows"); return 1; }
    prox_rows_kernel<<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, int(cols), mode, p, out, norms);
    return check_launch("prox_rows");
}

extern "C" int fb200_prox_apply(const double* x, int prox, double p0, double p1, int64_t n, double* out,
                                const double* scal, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid  = vec_grid(n);
    switch (prox) {
        case FB200_PROX_IDENTITY: prox_apply_kernel<FB200_PROX_IDENTITY><<<grid, VEC_THREADS, 0, st>>>(x, p0, p1, n, out, scal); break;
        case FB200_PROX_SHRINK:   prox_apply_kernel<FB200_PROX_SHRINK><<<grid, VEC_THREADS, 0, st>>>(x, p0, p1, n, out, scal); break;
        case FB200_PROX_NONNEG:   prox_apply_kernel<FB200_PROX_NONNEG><<<grid, VEC_THREADS, 0, st>>>(x, p0, p1, n, out, scal); break;
        case FB200_PROX_BOX:      prox_apply_kernel<FB200_PROX_BOX><<<grid, VEC_THREADS, 0, st>>>(x, p0, p1, n, out, scal); break;
        case FB200_PROX_L1BALL:   prox_apply_kernel<FB200_PROX_L1BALL><<<grid, VEC_THREADS, 0, st>>>(x, p0, p1, n, out, scal); break;
        case FB200_PROX_TV_BALL:
            if (n % 2) { set_error("prox_apply: TV ball prox needs an even element count"); return 1; }
            prox_apply_kernel<FB200_PROX_TV_BALL><<<grid, VEC_THREADS, 0, st>>>(x, p0, p1, n, out, scal);
            break;
        default: set_error("unknown prox tag %d", prox); return 1;
    }
    return check_launch("prox_apply");
}

extern "C" int fb200_l1ball_threshold(const double* v, int64_t n, double radius, double* scal, void* ws,
                                      void* stream) {
    (void)ws;
    l1ball_threshold_kernel<<<1, L1B_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(v, n, radius, scal);
    return check_launch("l1ball_threshold");
}

extern "C" int fb200_accel_step(double c, const double* xa1, const double* xa0, const double* xhat, int64_t n,
                                double* x1, const double* za1, const double* za0, const double* b, int64_t m,
                                int loss, int prox, double* z1, double* r, double* scal, void* ws, void* stream) {
    Workspace w(ws);
    cudaStream_t st  = static_cast<cudaStream_t>(stream);
    const int grid   = vec_grid(n > m ? n : m);
    const bool pen   = (prox == FB200_PROX_SHRINK);
#define FB200_ACCEL(LOSS)                                                                                        \
    if (pen) accel_step_kernel<LOSS, true><<<grid, VEC_THREADS, 0, st>>>(c, xa1, xa0, xhat, n, x1, za1, za0, b, m, z1, r, scal, w.red, w.counter); \
    else     accel_step_kernel<LOSS, false><<<grid, VEC_THREADS, 0, st>>>(c, xa1, xa0, xhat, n, x1, za1, za0, b, m, z1, r, scal, w.red, w.counter);
    switch (loss) {
        case FB200_LOSS_NONE:          FB200_ACCEL(FB200_LOSS_NONE) break;
        case FB200_LOSS_LEAST_SQUARES: FB200_ACCEL(FB200_LOSS_LEAST_SQUARES) break;
        case FB200_LOSS_LOGISTIC:      FB200_ACCEL(FB200_LOSS_LOGISTIC) break;
        default: set_error("unknown loss tag %d", loss); return 1;
    }
#undef FB200_ACCEL
    return check_launch("accel_step");
}

extern "C" int fb200_decide_init(double* scal, double f0, double g0_sq, void* stream) {
    if (!scal) { set_error("decide_init: null scalar block"); return 1; }
    decide_init_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(scal, f0, g0_sq);
    return check_launch("decide_init");
}

extern "C" int fb200_trial_decide(double* scal, double tau, int loss, int adaptive, int backtrack, int bt,
                                  int max_backtracks, int window, int stop_rule, double tolerance, int host_it,
                                  double host_max_residual, double host_g0_sq, double* host_out, void* stream) {
    if (!scal) { set_error("trial_decide: null scalar block"); return 1; }
    if (window < 1 || window > FB200_FRING) { set_error("trial_decide: window %d not in 1..%d", window, FB200_FRING); return 1; }
    DecideArgs d{};
    d.loss = loss; d.adaptive = adaptive; d.backtrack = backtrack; d.bt = bt; d.max_backtracks = max_backtracks;
    d.window = window; d.stop_rule = stop_rule; d.host_it = host_it;
    d.tolerance = tolerance; d.host_max_residual = host_max_residual; d.host_g0_sq = host_g0_sq;
    trial_decide_kernel<<<1, FB200_NSCAL, 0, static_cast<cudaStream_t>(stream)>>>(scal, tau, d, host_out);
    return check_launch("trial_decide");
}

extern "C" int fb200_loss_eval(int loss, const double* z, const double* b, int64_t m, double* r, double* scal,
                               void* ws, void* stream) {
    Workspace w(ws);
    return launch_loss(loss, z, 0, 0, b, m, nullptr, r, scal, w, static_cast<cudaStream_t>(stream));
}

extern "C" int fb200_bb_reduce(const double* g1, const double* x0, const double* xhat, const double* dx,
                               double tau, int64_t n, int adaptive, double* scal, void* ws, void* stream) {
    Workspace w(ws);
    return launch_bb(adaptive ? 2 : 1, g1, 0, 0, n, nullptr, x0, xhat, dx, tau, scal, w,
                     static_cast<cudaStream_t>(stream));
}

template <int OP>
static int launch_reduce(const double* a, const double* b, int64_t n, double* out, void* ws, void* stream,
                         const char* name) {
    Workspace w(ws);
    reduce_kernel<OP><<<vec_grid(n), VEC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out, w.red,
                                                                                         w.counter);
    return check_launch(name);
}

extern "C" int fb200_dot(const double* a, const double* b, int64_t n, double* out, void* ws, void* stream) {
    return launch_reduce<0>(a, b, n, out, ws, stream, "dot");
}
extern "C" int fb200_diff_nrm2sq(const double* a, const double* b, int64_t n, double* out, void* ws, void* stream) {
    return launch_reduce<1>(a, b, n, out, ws, stream, "diff_nrm2sq");
}
extern "C" int fb200_asum(const double* a, int64_t n, double* out, void* ws, void* stream) {
    return launch_reduce<2>(a, nullptr, n, out, ws, stream, "asum");
}
extern "C" int fb200_amax(const double* a, int64_t n, double* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    amax_zero_kernel<<<1, 1, 0, st>>>(out);
    if (n > 0) amax_kernel<<<vec_grid(n), VEC_THREADS, 0, st>>>(a, n, out);
    return check_launch("amax");
}

// Device-resident FBS loop for SMALL dense problems (A fits the L2: BASELINE config 1, the reference's own
// 200 x 1000 lasso and its relatives).
//
// At this size an iteration of the host-driven path is ~100 us of launch, sync and Python latency around
// ~5 us of arithmetic.  Here the WHOLE loop of the reference (fasta/__init__.py:172-313: forward step, prox,
// A x, f, non-monotone backtracking line search, A^T r, Barzilai-Borwein step size, residuals, best iterate,
// stop rule) runs inside ONE cooperative kernel; the phases of an iteration are separated by grid-wide
// barriers, every block forms every cross-block sum in the same fixed order and therefore takes the same
// scalar decisions, and the histories are written to device arrays that the host reads once at the end.
// All three modes (plain, adaptive, FISTA), built-in stop rules, elementwise prox (shrink / nonneg / box / identity).
//
// Compiled with -fmad=false: the elementwise lines and the scalar step-size algebra round once per numpy
// operation of the reference line (np.float64 scalars are IEEE doubles); dot products use explicit fma().
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace fb200 {

constexpr int RL_THREADS = 256;
constexpr int RL_MAXK    = 5;

struct ResidentArgs {
    const double* A; int64_t lda; int M, N;
    const double* b;
    double* X[2]; double* G[2];          // ping-pong iterate / gradient; [0] holds the start point and its gradient
    double* XA[2]; double* ZA[2];        // FISTA: prox points and their images; [0] = start point and A x0
    double *xhat, *dx, *best, *z, *r;
    double* part;                        // [6][grid][RL_MAXK] per-block partial sums (two copies per phase)
    double *resid_h, *nresid_h, *tau_h, *f_h, *obj_h;   // histories; f_h[0] / obj_h[0] preset by the host
    int* bt_h;                           // backtracks per iteration; bit 30 set = acceleration restarted
    double* alpha_h;                     // FISTA: alpha0 per iteration (verbose line)
    unsigned long long* clock_h;         // %globaltimer at the start of every iteration and at exit
    double* out;                         // [0] iterations, [1] total backtracks, [2] which buffer holds the last iterate
    double tau_init, g1_sq_init, tolerance, shrink, pen_mu, p_lo, p_hi;
    int adaptive, backtrack, window, max_backtracks, max_iters, stop_rule, evaluate_objective, restart;
};

__device__ __forceinline__ unsigned long long rl_clock() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// sum of the per-block partials in block order, identical in every block; result broadcast to all threads
template <int K>
__device__ __forceinline__ void rl_collect(const double* part, int nblocks, double (&tot)[K], double* sm) {
    if (threadIdx.x < 32) {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int bk = threadIdx.x; bk < nblocks; bk += 32) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __ldcg(&part[bk * RL_MAXK + k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            acc[k] = warp_sum(acc[k]);
            if (threadIdx.x == 0) sm[k] = acc[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = sm[k];
    __syncthreads();
}

template <int K>
__device__ __forceinline__ void rl_publish(double (&v)[K], double* part, double* sm) {
    block_sum<K>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) part[blockIdx.x * RL_MAXK + k] = v[k];
    }
}

__device__ __forceinline__ double rl_sq(double v) {          // la.norm(.)**2 = sqrt(dot)**2
    const double t = sqrt(v);
    return t * t;
}
__device__ __forceinline__ double rl_pymax(double a, double b) { return (b > a) ? b : a; }   // Python max(a, b)

// CLUSTER variant (problems whose matrix fits the shared memory of one 16-CTA cluster TWICE -- BASELINE config 1 does):
// the grid is a single thread-block cluster; every CTA keeps its band of rows of A (for A x) and its 32-column groups of
// A (for A^T r) in shared memory, so after the one-time staging no phase touches A in L2 / HBM again, and the phases
// are separated by the hardware cluster barrier (release / acquire at cluster scope) instead of a grid-wide barrier.
// Same per-row and per-column summation order as the grid variant: identical z and g, bit for bit.
constexpr int RL_CLUSTER = 16;

template <bool CLUSTER>
__device__ __forceinline__ void rl_barrier(cg::grid_group& grid) {
    if (CLUSTER) cg::this_cluster().sync();
    else grid.sync();
}

template <int LOSS, int PROX, bool ACCEL, bool CLUSTER>
__global__ void __launch_bounds__(RL_THREADS)
resident_fbs_kernel(ResidentArgs p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[RL_MAXK * 32 + RL_MAXK];
    __shared__ double gsm[8][33];
    extern __shared__ double rl_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = gridDim.x;
    const int gtid = blockIdx.x * RL_THREADS + tid, gthreads = nb * RL_THREADS;
    const int gwarp = blockIdx.x * (RL_THREADS / 32) + warp, gwarps = nb * (RL_THREADS / 32);
    const int M = p.M, N = p.N;
    // CLUSTER: rows [row_lo, row_hi) of A as rows_sm[row - row_lo][N]; column groups as cols_sm[row][grp][32]
    const int rpb = (M + nb - 1) / nb, ngrp = (N + nb * 32 - 1) / (nb * 32);
    const int row_lo = min(M, int(blockIdx.x) * rpb), row_hi = min(M, row_lo + rpb);
    double* rows_sm = rl_dyn;
    double* cols_sm = rl_dyn + ((size_t(rpb) * N + 1) & ~size_t(1));
    double* xv_sm = cols_sm + size_t(M) * ngrp * 32;     // [N] the vector of the current A x      (staged from L2 once per phase)
    double* rv_sm = xv_sm + ((N + 1) & ~1);              // [M] the vector of the current A^T r
    if (CLUSTER) {
        for (int e = tid; e < (row_hi - row_lo) * N; e += RL_THREADS) {
            const int rr = e / N, cc = e - rr * N;
            rows_sm[e] = p.A[int64_t(row_lo + rr) * p.lda + cc];
        }
        for (int e = tid; e < M * ngrp * 32; e += RL_THREADS) {
            const int rr = e / (ngrp * 32), rem = e - rr * (ngrp * 32), grp = rem >> 5;
            const int col = int(blockIdx.x) * 32 + grp * nb * 32 + (rem & 31);
            cols_sm[e] = (col < N) ? p.A[int64_t(rr) * p.lda + col] : 0.0;
        }
        __syncthreads();
    }
    double* partA[2] = {p.part, p.part + size_t(nb) * RL_MAXK};
    double* partB[2] = {p.part + 2 * size_t(nb) * RL_MAXK, p.part + 3 * size_t(nb) * RL_MAXK};
    double* partC[2] = {p.part + 4 * size_t(nb) * RL_MAXK, p.part + 5 * size_t(nb) * RL_MAXK};
    double* partE[2] = {p.part + 6 * size_t(nb) * RL_MAXK, p.part + 7 * size_t(nb) * RL_MAXK};
    unsigned ua = 0, ub = 0, uc = 0, ue = 0;     // uses of each partial buffer (toggle the copy)
    double alpha1 = 1.0;                 // reference :157
    int acur = 0;                        // XA[acur] / ZA[acur]: previous prox point and its image

    // the last 64 values of f_hist, kept per block (every block computes the same f1): the non-monotone window needs no
    // grid-wide visibility of f_h, which saves the fourth barrier of an iteration (window > 64: global history + barrier)
    __shared__ double fwin[64];
    const bool local_win = p.window <= 64;
    if (tid == 0) fwin[0] = p.f_h[0];
    __syncthreads();
    double tau1 = p.tau_init, g1_sq = p.g1_sq_init;
    double max_residual = -INFINITY, best_q = INFINITY;
    int cur = 0, it = 0;
    long long total_bt = 0;

    while (it < p.max_iters) {
        if (gtid == 0) p.clock_h[it] = rl_clock();
        const double* x0 = p.X[cur];
        const double* g0 = p.G[cur];
        double* x1 = p.X[1 - cur];
        double* g1 = p.G[1 - cur];
        const double g0_sq = g1_sq;
        double tau0 = tau1;
        int bt = 0;
        double f_window_max = -INFINITY;
        for (int k = (it - p.window + 1 > 0 ? it - p.window + 1 : 0); k <= it; ++k)
            f_window_max = fmax(f_window_max, local_win ? fwin[k & 63] : __ldcg(&p.f_h[k]));
        double dx_g0, dx_sq, xmxh_sq, pen_raw, f1, restart_dot = 0.0;
        while (true) {
            // ---- forward step, prox, Dx and their sums (reference :181-186) ----
            double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            const double p0 = (PROX == FB200_PROX_SHRINK) ? tau0 * p.pen_mu : p.p_lo;
            double* xp = ACCEL ? p.XA[1 - acur] : x1;             // the prox point (x_accel1 when accelerating)
            for (int i = gtid; i < N; i += gthreads) {
                const double a = __ldcg(&x0[i]), gr = __ldcg(&g0[i]);
                const double h = a - tau0 * gr;
                const double y = prox_elem<PROX>(h, p0, p.p_hi);
                const double d = y - a;
                p.xhat[i] = h;
                xp[i] = y;
                p.dx[i] = d;
                s[0] += d * gr;
                s[1] += d * d;
                const double e = y - h;
                s[2] += e * e;
                s[3] += fabs(y);
                if (ACCEL) s[4] += (a - y) * (y - __ldcg(&p.XA[acur][i]));      // restart test, reference :231
            }
            rl_publish<5>(s, partA[ua & 1], sm);
            rl_barrier<CLUSTER>(grid);
            // ---- z = A x1, r = gradf(z), f (reference :187-188): one warp per row ----
            double fs[1] = {0.0};
            double* zp = ACCEL ? p.ZA[1 - acur] : p.z;
            if (CLUSTER) {
                for (int j = tid; j < N; j += RL_THREADS) xv_sm[j] = __ldcg(&xp[j]);
                __syncthreads();
            }
            for (int row = CLUSTER ? row_lo + warp : gwarp; row < (CLUSTER ? row_hi : M); row += CLUSTER ? RL_THREADS / 32 : gwarps) {
                const double* ar = CLUSTER ? rows_sm + size_t(row - row_lo) * N : p.A + int64_t(row) * p.lda;
                double acc = 0.0;
                if (CLUSTER)
                    for (int j = lane; j < N; j += 32) acc = fma(ar[j], xv_sm[j], acc);
                else
                    for (int j = lane; j < N; j += 32) acc = fma(ar[j], __ldcg(&xp[j]), acc);
                acc = warp_sum(acc);
                if (lane == 0) {
                    double ri, fi;
                    loss_elem<LOSS>(acc, p.b[row], ri, fi);
                    zp[row] = acc;
                    if (!ACCEL) p.r[row] = ri;
                    fs[0] += fi;
                }
            }
            rl_publish<1>(fs, partB[ub & 1], sm);
            rl_barrier<CLUSTER>(grid);
            double ta[5], tb[1];
            rl_collect<5>(partA[ua & 1], nb, ta, sm);
            rl_collect<1>(partB[ub & 1], nb, tb, sm);
            ++ua; ++ub;
            dx_g0 = ta[0]; dx_sq = ta[1]; xmxh_sq = ta[2]; pen_raw = ta[3]; restart_dot = ta[4];
            f1 = (LOSS == FB200_LOSS_LEAST_SQUARES) ? .5 * rl_sq(tb[0]) : tb[0];
            // ---- non-monotone line search (reference :195-217) ----
            if (p.backtrack && (f1 - (f_window_max + dx_g0 + rl_sq(dx_sq) / (2 * tau0)) > 1E-12) && bt < p.max_backtracks) {
                tau0 *= p.shrink;
                ++bt;
                continue;
            }
            break;
        }
        total_bt += bt;
        double alpha0 = 0.0;
        bool restarted = false;
        if (ACCEL) {
            // ---- FISTA extrapolation of x and z, f at the extrapolated z (reference :220-245) ----
            alpha0 = alpha1;
            if (p.restart && restart_dot > 1E-30) { alpha0 = 1.0; restarted = true; }
            alpha1 = (1 + sqrt(1 + 4 * (alpha0 * alpha0))) / 2;
            const double c = (alpha0 - 1) / alpha1;
            const double* xa1 = p.XA[1 - acur];
            const double* xa0 = p.XA[acur];
            const double* za1 = p.ZA[1 - acur];
            const double* za0 = p.ZA[acur];
            double se[3] = {0.0, 0.0, 0.0};
            for (int i = gtid; i < N; i += gthreads) {
                const double q = __ldcg(&xa1[i]);
                const double y = q + c * (q - __ldcg(&xa0[i]));
                x1[i] = y;
                const double e = y - __ldcg(&p.xhat[i]);
                se[1] += e * e;
                se[2] += fabs(y);
            }
            for (int row = gtid; row < M; row += gthreads) {
                const double q = __ldcg(&za1[row]);
                const double zz = q + c * (q - __ldcg(&za0[row]));
                double ri, fi;
                loss_elem<LOSS>(zz, p.b[row], ri, fi);
                p.z[row] = zz;
                p.r[row] = ri;
                se[0] += fi;
            }
            rl_publish<3>(se, partE[ue & 1], sm);
            rl_barrier<CLUSTER>(grid);
            double te[3];
            rl_collect<3>(partE[ue & 1], nb, te, sm);
            ++ue;
            f1 = (LOSS == FB200_LOSS_LEAST_SQUARES) ? .5 * rl_sq(te[0]) : te[0];
            xmxh_sq = te[1];
            pen_raw = te[2];
            acur = 1 - acur;
        }
        // ---- g1 = A^T r (reference :248): 32 columns x 8 row lanes per block pass, + BB sums (:254-260) ----
        double sc[3] = {0.0, 0.0, 0.0};
        int grp = 0;
        if (CLUSTER) {
            for (int j = tid; j < M; j += RL_THREADS) rv_sm[j] = __ldcg(&p.r[j]);
            __syncthreads();
        }
        for (int c0 = blockIdx.x * 32; c0 < N; c0 += nb * 32, ++grp) {
            const int col = c0 + lane;
            double acc = 0.0;
            if (col < N) {
                if (CLUSTER)
                    for (int row = warp; row < M; row += 8) acc = fma(cols_sm[(size_t(row) * ngrp + grp) * 32 + lane], rv_sm[row], acc);
                else
                    for (int row = warp; row < M; row += 8) acc = fma(p.A[int64_t(row) * p.lda + col], __ldcg(&p.r[row]), acc);
            }
            gsm[warp][lane] = acc;
            __syncthreads();
            if (warp == 0 && col < N) {
                double gi = gsm[0][lane];
#pragma unroll
                for (int w = 1; w < 8; ++w) gi += gsm[w][lane];
                g1[col] = gi;
                sc[2] += gi * gi;
                if (p.adaptive) {
                    const double dg = gi + (__ldcg(&p.xhat[col]) - __ldcg(&x0[col])) / tau0;
                    sc[0] += __ldcg(&p.dx[col]) * dg;
                    sc[1] += dg * dg;
                }
            }
            __syncthreads();
        }
        rl_publish<3>(sc, partC[uc & 1], sm);
        rl_barrier<CLUSTER>(grid);
        double tc[3];
        rl_collect<3>(partC[uc & 1], nb, tc, sm);
        ++uc;
        g1_sq = tc[2];
        // ---- step-size algebra, residuals, histories (reference :253-300) ----
        const double dx_norm = sqrt(dx_sq);
        tau1 = tau0;
        if (p.adaptive) {
            const double dotprod = tc[0];
            const double tau_s = (dx_norm * dx_norm) / dotprod;
            const double q = dotprod / rl_sq(tc[1]);
            const double tau_m = (0.0 > q) ? 0.0 : q;             // Python max(q, 0): a nan q stays
            if (2 * tau_m > tau_s) tau1 = tau_m;
            else tau1 = tau_s - .5 * tau_m;
            if (tau1 <= 0 || isinf(tau1) || isnan(tau1)) tau1 = tau0 * 1.5;
        }
        const double resid = dx_norm / tau0;
        const double normalizer = rl_pymax(sqrt(g0_sq), sqrt(xmxh_sq) / tau0) + 1E-12;
        const double nresid = resid / normalizer;
        max_residual = rl_pymax(max_residual, resid);
        const double objective = f1 + ((PROX == FB200_PROX_SHRINK) ? p.pen_mu * pen_raw : 0.0);
        const double quality = p.evaluate_objective ? objective : resid;
        if (gtid == 0) {
            p.resid_h[it] = resid;
            p.nresid_h[it] = nresid;
            p.tau_h[it] = tau0;
            p.f_h[it + 1] = f1;
            if (p.evaluate_objective) p.obj_h[it + 1] = objective;
            p.bt_h[it] = bt | (restarted ? (1 << 30) : 0);
            if (ACCEL) p.alpha_h[it] = alpha0;
        }
        if (quality < best_q) {
            for (int i = gtid; i < N; i += gthreads) p.best[i] = __ldcg(&x1[i]);
            best_q = quality;
        }
        bool stop;
        switch (p.stop_rule) {
            case 0: stop = resid < p.tolerance; break;
            case 1: stop = nresid < p.tolerance; break;
            case 2: stop = resid / max_residual < p.tolerance; break;
            default: stop = (resid / max_residual < p.tolerance) || (nresid < p.tolerance); break;
        }
        cur = 1 - cur;
        ++it;
        // f_h[it] must be visible to every block before the next window maximum (kept per block when the window fits
        // the local ring); the partial buffers and xhat / dx / z / r are protected by the two barriers of the next trial
        if (local_win) {
            if (tid == 0) fwin[it & 63] = f1;       // it was incremented: this is f_hist[it]
            __syncthreads();
        } else {
            rl_barrier<CLUSTER>(grid);
        }
        if (stop) break;
    }
    if (gtid == 0) {
        p.clock_h[it] = rl_clock();
        p.out[0] = double(it);
        p.out[1] = double(total_bt);
        p.out[2] = double(cur);
        p.out[3] = CLUSTER ? 1.0 : 0.0;
    }
}

typedef void (*ResidentKernel)(ResidentArgs);

template <int LOSS, bool ACCEL, bool CLUSTER>
static ResidentKernel resident_pick(int prox) {
    switch (prox) {
        case FB200_PROX_SHRINK: return resident_fbs_kernel<LOSS, FB200_PROX_SHRINK, ACCEL, CLUSTER>;
        case FB200_PROX_NONNEG: return resident_fbs_kernel<LOSS, FB200_PROX_NONNEG, ACCEL, CLUSTER>;
        case FB200_PROX_BOX: return resident_fbs_kernel<LOSS, FB200_PROX_BOX, ACCEL, CLUSTER>;
        case FB200_PROX_IDENTITY: return resident_fbs_kernel<LOSS, FB200_PROX_IDENTITY, ACCEL, CLUSTER>;
        default: return nullptr;
    }
}

template <bool CLUSTER>
static ResidentKernel resident_kernel(int loss, int prox, int accelerate) {
    if (loss == FB200_LOSS_LEAST_SQUARES)
        return accelerate ? resident_pick<FB200_LOSS_LEAST_SQUARES, true, CLUSTER>(prox) : resident_pick<FB200_LOSS_LEAST_SQUARES, false, CLUSTER>(prox);
    if (loss == FB200_LOSS_LOGISTIC)
        return accelerate ? resident_pick<FB200_LOSS_LOGISTIC, true, CLUSTER>(prox) : resident_pick<FB200_LOSS_LOGISTIC, false, CLUSTER>(prox);
    return nullptr;
}

// dynamic shared memory of the cluster variant (0 = the matrix does not fit twice into one cluster's shared memory)
static size_t resident_cluster_smem(int64_t M, int64_t N) {
    const int64_t rpb = (M + RL_CLUSTER - 1) / RL_CLUSTER, ngrp = (N + RL_CLUSTER * 32 - 1) / (RL_CLUSTER * 32);
    const int64_t doubles = ((rpb * N + 1) & ~int64_t(1)) + M * ngrp * 32 + ((N + 1) & ~int64_t(1)) + M;
    const int64_t bytes = doubles * 8;
    return bytes <= 220 * 1024 ? size_t(bytes) : 0;      // + 3.4 KB static, of the 227 KB a CTA may use
}

}  // namespace fb200

using namespace fb200;

// blocks the resident loop would use for an M x N problem (0 = not eligible: too large for the L2, or the
// device cannot co-schedule the grid)
extern "C" int fb200_resident_blocks(int64_t M, int64_t N) {
    if (M < 1 || N < 1 || M > (1 << 24) || N > (1 << 24) || M * N * 8 > (int64_t(48) << 20)) return 0;
    int64_t want = (M + 7) / 8;                       // one warp per row
    if ((N + 31) / 32 > want) want = (N + 31) / 32;   // 32 columns per block pass
    const int cap = sm_count();
    return int(want < cap ? want : cap);
}

extern "C" size_t fb200_resident_scratch_doubles(int64_t M, int64_t N) {
    const int nb = fb200_resident_blocks(M, N);
    return size_t(8) * size_t(nb > RL_CLUSTER ? nb : RL_CLUSTER) * RL_MAXK;
}

// 1 if the single-cluster variant (matrix resident in the cluster's shared memory) can run this problem
extern "C" int fb200_resident_cluster_ok(int64_t M, int64_t N) {
    if (fb200_resident_blocks(M, N) < 1 || resident_cluster_smem(M, N) == 0) return 0;
    const char* e = getenv("FASTA_B200_RESIDENT_CLUSTER");
    return (e && e[0] == '0') ? 0 : 1;
}

// One launch = the whole solve after the prologue.  Pointers as in ResidentArgs; stop_rule 0..3 = residual,
// norm_residual, ratio_residual, hybrid_residual (reference stopping.py:15,27,39,51).
extern "C" int fb200_resident_fbs(const double* A, int64_t lda, int64_t M, int64_t N, const double* b, int loss, int prox,
                                  double pen_mu, double p_lo, double p_hi, double* x_a, double* x_b, double* g_a, double* g_b,
                                  double* xhat, double* dx, double* best, double* z, double* r, double* part,
                                  double* resid_h, double* nresid_h, double* tau_h, double* f_h, double* obj_h, int* bt_h,
                                  unsigned long long* clock_h, double* out, double tau_init, double g1_sq_init,
                                  double tolerance, double shrink, int adaptive, int backtrack, int window,
                                  int max_backtracks, int max_iters, int stop_rule, int evaluate_objective, int accelerate,
                                  int restart, double* xa_a, double* xa_b, double* za_a, double* za_b, double* alpha_h,
                                  void* stream) {
    const int nb = fb200_resident_blocks(M, N);
    if (nb < 1) { set_error("resident_fbs: problem not eligible"); return 1; }
    if (accelerate && (!xa_a || !xa_b || !za_a || !za_b || !alpha_h)) { set_error("resident_fbs: FISTA buffers missing"); return 1; }
    ResidentKernel k = resident_kernel<false>(loss, prox, accelerate);
    if (!k) { set_error("resident_fbs: unsupported loss / prox tags %d / %d", loss, prox); return 1; }
    // single-cluster variant: the matrix stays in the cluster's shared memory
    ResidentKernel kc = nullptr;
    size_t csmem = 0;
    if (fb200_resident_cluster_ok(M, N)) {
        kc = resident_kernel<true>(loss, prox, accelerate);
        csmem = resident_cluster_smem(M, N);
        // attributes and the placement query once per kernel and shared-memory size
        // (per device: attributes and placement belong to the current device's context)
        static std::mutex ok_mutex;
        static ResidentKernel ok_kernel[64];
        static size_t ok_smem[64];
        static int ok_state[64], ok_dev[64], ok_count = 0;
        const int dev = current_device();
        std::lock_guard<std::mutex> ok_lock(ok_mutex);
        int state = -1;
        for (int i = 0; i < ok_count; ++i)
            if (ok_kernel[i] == kc && ok_smem[i] == csmem && ok_dev[i] == dev) state = ok_state[i];
        if (state < 0) {
            cudaLaunchConfig_t probe{};
            cudaLaunchAttribute pattr[1];
            probe.gridDim = dim3(RL_CLUSTER); probe.blockDim = dim3(RL_THREADS); probe.dynamicSmemBytes = csmem;
            pattr[0].id = cudaLaunchAttributeClusterDimension;
            pattr[0].val.clusterDim.x = RL_CLUSTER; pattr[0].val.clusterDim.y = 1; pattr[0].val.clusterDim.z = 1;
            probe.attrs = pattr; probe.numAttrs = 1;
            int nclusters = 0;
            state = 1;
            if (cudaFuncSetAttribute(reinterpret_cast<const void*>(kc), cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                cudaFuncSetAttribute(reinterpret_cast<const void*>(kc), cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
                cudaOccupancyMaxActiveClusters(&nclusters, reinterpret_cast<const void*>(kc), &probe) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                state = 0;                              // this device cannot place the cluster: grid variant
            }
            if (ok_count < 64) { ok_kernel[ok_count] = kc; ok_smem[ok_count] = csmem; ok_state[ok_count] = state; ok_dev[ok_count] = dev; ++ok_count; }
        }
        if (!state) kc = nullptr;
    }
    int per_sm = 0;
    if (!kc && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, RL_THREADS, 0) != cudaSuccess || per_sm < 1)) {
        cudaGetLastError();
        set_error("resident_fbs: occupancy query failed");
        return 1;
    }
    ResidentArgs a{};
    a.A = A; a.lda = lda; a.M = int(M); a.N = int(N); a.b = b;
    a.X[0] = x_a; a.X[1] = x_b; a.G[0] = g_a; a.G[1] = g_b;
    a.xhat = xhat; a.dx = dx; a.best = best; a.z = z; a.r = r; a.part = part;
    a.resid_h = resid_h; a.nresid_h = nresid_h; a.tau_h = tau_h; a.f_h = f_h; a.obj_h = obj_h; a.bt_h = bt_h;
    a.clock_h = clock_h; a.out = out;
    a.XA[0] = xa_a; a.XA[1] = xa_b; a.ZA[0] = za_a; a.ZA[1] = za_b; a.alpha_h = alpha_h; a.restart = restart;
    a.tau_init = tau_init; a.g1_sq_init = g1_sq_init; a.tolerance = tolerance; a.shrink = shrink;
    a.pen_mu = pen_mu; a.p_lo = p_lo; a.p_hi = p_hi;
    a.adaptive = adaptive; a.backtrack = backtrack; a.window = window; a.max_backtracks = max_backtracks;
    a.max_iters = max_iters; a.stop_rule = stop_rule; a.evaluate_objective = evaluate_objective;
    if (kc) {
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        cfg.gridDim = dim3(RL_CLUSTER); cfg.blockDim = dim3(RL_THREADS); cfg.dynamicSmemBytes = csmem;
        cfg.stream = static_cast<cudaStream_t>(stream);
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = RL_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t ec = cudaLaunchKernelEx(&cfg, kc, a);
        if (ec != cudaSuccess) { set_error("resident_fbs: cluster launch failed: %s", cudaGetErrorString(ec)); cudaGetLastError(); return 1; }
        return 0;
    }
    void* params[1] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k), dim3(unsigned(nb)), dim3(RL_THREADS), params, 0,
                                                static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { set_error("resident_fbs: cooperative launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    return 0;
}

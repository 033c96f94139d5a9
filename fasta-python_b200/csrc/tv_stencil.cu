// K11 / K12: matrix-free total-variation operators (periodic finite differences) fused with the
// loss and Barzilai-Borwein epilogues.
//     div : Y (n0 x n1 x 2) -> Z (n0 x n1)     reference tv_denoising.py:43-63   (solver map A)
//     grad: R (n0 x n1) -> G (n0 x n1 x 2)     reference tv_denoising.py:26-40   (adjoint A^H)
//
// B200 design: HBM-bound stencils.  A thread owns one image column inside a strip of STRIP rows
// and marches down it, so the vertical neighbour is carried in registers (each input element is
// loaded from HBM once per strip, +1 halo row per strip = 3% re-read); the horizontal neighbour
// comes from the adjacent lane by warp shuffle (one extra L1-hit load per warp edge).  Loads are
// issued UNROLL rows ahead to keep >100 KB per SM in flight.  Outputs, the loss (r = gradf(z),
// sum f) and the BB reductions are produced in the same pass, so an FBS iteration on an image
// is three streaming kernels.  Compiled with -fmad=false (one rounding per numpy operation).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fb200 {

constexpr int TV_THREADS = 128;
constexpr int TV_STRIP   = 32;   // minimum rows per strip; grows so the grid fits the reduction workspace
constexpr int TV_UNROLL  = 4;

// Z[i][j] = (Y[i+1][j].x - Y[i][j].x) + (Y[i][j+1].y - Y[i][j].y), indices periodic
template <int LOSS>
__global__ void __launch_bounds__(TV_THREADS)
tv_div_loss_kernel(const double2* __restrict__ Y, int64_t n0, int64_t n1, const double* __restrict__ b,
                   double* __restrict__ z, double* __restrict__ r, int strip, double* scal, double* red,
                   unsigned* counter) {
    const int lane   = threadIdx.x & 31;
    const int64_t j  = int64_t(blockIdx.x) * TV_THREADS + threadIdx.x;
    const int64_t i0 = int64_t(blockIdx.y) * strip;
    const bool live  = j < n1;
    const int64_t jc = live ? j : 0;
    const int64_t jr = (jc + 1 == n1) ? 0 : jc + 1;
    double s[1] = {0.0};
    if (i0 < n0) {
        const int64_t i1 = (i0 + strip < n0) ? i0 + strip : n0;
        double2 cur = Y[i0 * n1 + jc];
        for (int64_t ib = i0; ib < i1; ib += TV_UNROLL) {
            double2 nxt[TV_UNROLL];
            double  ry[TV_UNROLL];
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i  = ib + u;
                const int64_t in = (i + 1 >= n0) ? (i + 1 - n0) : i + 1;
                nxt[u] = (i < i1) ? Y[in * n1 + jc] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i  = ib + u;
                const double2 c  = (u == 0) ? cur : nxt[u - 1];
                double right     = __shfl_down_sync(0xffffffffu, c.y, 1);
                if ((lane == 31 || jc + 1 == n1) && i < i1) right = Y[i * n1 + jr].y;
                ry[u] = right;
            }
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                const double2 c = (u == 0) ? cur : nxt[u - 1];
                if (i < i1 && live) {
                    const double t0 = nxt[u].x - c.x;
                    const double t1 = ry[u] - c.y;
                    const double zi = t0 + t1;
                    const int64_t o = i * n1 + j;
                    z[o] = zi;
                    if (LOSS != FB200_LOSS_NONE) {
                        double ri, fi;
                        loss_elem<LOSS>(zi, b[o], ri, fi);
                        r[o] = ri;
                        s[0] += fi;
                    }
                }
            }
            cur = nxt[TV_UNROLL - 1];
        }
    }
    if (LOSS != FB200_LOSS_NONE) {
        double* const out[1] = {scal + FB200_S_F};
        grid_sum<1>(s, red, counter, out);
    }
}

// G[i][j] = (R[i-1][j] - R[i][j], R[i][j-1] - R[i][j]), indices periodic; fused BB reductions
template <int BB>
__global__ void __launch_bounds__(TV_THREADS)
tv_grad_bb_kernel(const double* __restrict__ R, int64_t n0, int64_t n1, double2* __restrict__ g,
                  const double2* __restrict__ x0, const double2* __restrict__ xhat, const double2* __restrict__ dx,
                  double tau, int strip, double* scal, double* red, unsigned* counter) {
    const int lane   = threadIdx.x & 31;
    const int64_t j  = int64_t(blockIdx.x) * TV_THREADS + threadIdx.x;
    const int64_t i0 = int64_t(blockIdx.y) * strip;
    const bool live  = j < n1;
    const int64_t jc = live ? j : 0;
    const int64_t jl = (jc == 0) ? n1 - 1 : jc - 1;
    double s[3] = {0.0, 0.0, 0.0};
    if (i0 < n0) {
        const int64_t i1 = (i0 + strip < n0) ? i0 + strip : n0;
        double up = R[((i0 == 0) ? n0 - 1 : i0 - 1) * n1 + jc];
        for (int64_t ib = i0; ib < i1; ib += TV_UNROLL) {
            double c[TV_UNROLL], lf[TV_UNROLL];
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                c[u] = (i < i1) ? R[i * n1 + jc] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                double left = __shfl_up_sync(0xffffffffu, c[u], 1);
                if ((lane == 0 || jc == 0) && i < i1) left = R[i * n1 + jl];
                lf[u] = left;
            }
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                if (i < i1 && live) {
                    const double a = (u == 0) ? up : c[u - 1];
                    double2 gi;
                    gi.x = a - c[u];
                    gi.y = lf[u] - c[u];
                    const int64_t o = i * n1 + j;
                    g[o] = gi;
                    if (BB >= 1) {
                        s[2] += gi.x * gi.x;
                        s[2] += gi.y * gi.y;
                    }
                    if (BB >= 2) {
                        const double2 h = xhat[o], p = x0[o], d = dx[o];
                        const double dg0 = gi.x + (h.x - p.x) / tau;
                        const double dg1 = gi.y + (h.y - p.y) / tau;
                        s[0] += d.x * dg0;
                        s[0] += d.y * dg1;
                        s[1] += dg0 * dg0;
                        s[1] += dg1 * dg1;
                    }
                }
            }
            up = c[TV_UNROLL - 1];
        }
    }
    if (BB >= 1) {
        double* const out[3] = {BB >= 2 ? scal + FB200_S_DX_DG : nullptr,
                                BB >= 2 ? scal + FB200_S_DG_SQ : nullptr, scal + FB200_S_G1_SQ};
        grid_sum<3>(s, red, counter, out);
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused TV iteration (non-accelerated modes): the forward step, the per-pixel ball projection, the
// divergence and the loss in ONE pass,
//     xhat = x0 - tau*g0 ; x1 = xhat / max(|xhat|_2, 1) ; z = div(x1) ; r = z - b ; sums
// (reference __init__.py:181-188 with tv_denoising.py:43-63,85-96).  xhat, dx and z are never
// written: the gradient kernel below recomputes xhat and dx from x0, g0, x1 with the same
// expressions (bit-identical), so an iteration moves 8U + 9U (adaptive) or 8U + 3U (plain) bytes
// instead of 24U / 18U, U = n0*n1*8.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 tv_prox_point(double2 a, double2 gr, double tau, double2& h) {
    h.x = a.x - tau * gr.x;
    h.y = a.y - tau * gr.y;
    const double nrm = fmax(sqrt(h.x * h.x + h.y * h.y), 1.0);
    return make_double2(h.x / nrm, h.y / nrm);
}

// Shared-memory stencil: a block walks 16 x 64-pixel tiles; for every pixel of the tile plus one halo
// row and one halo column it computes the prox point once into shared memory (so the sqrt / divide
// chain of a pixel is independent of its neighbours: no serial marching, full occupancy, every load
// independent), then forms the divergence and the loss from shared memory.
constexpr int TVS_TH = 16, TVS_TW = 64, TVS_THREADS = 256;
constexpr int TVS_EXT = (TVS_TH + 1) * (TVS_TW + 1);        // 1105 prox points per tile (8% halo)

template <int LOSS>
__global__ void __launch_bounds__(TVS_THREADS)
tv_step_div_loss_kernel(const double2* __restrict__ x0, const double2* __restrict__ g0, double tau, int64_t n0, int64_t n1,
                        const double* __restrict__ b, double2* __restrict__ x1, double* __restrict__ r, int tiles_x,
                        int ntiles, double* scal, double* red, unsigned* counter) {
    __shared__ double2 ys[TVS_EXT];
    double s[4] = {0.0, 0.0, 0.0, 0.0};      // <dx,g0>, <dx,dx>, |x1-xhat|^2, f
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = int64_t(tile / tiles_x) * TVS_TH, j0 = int64_t(tile % tiles_x) * TVS_TW;
        __syncthreads();                     // previous tile's reads of ys are done
        for (int e = threadIdx.x; e < TVS_EXT; e += TVS_THREADS) {
            const int ti = e / (TVS_TW + 1), tj = e - ti * (TVS_TW + 1);
            int64_t i = i0 + ti, j = j0 + tj;
            const bool mine = ti < TVS_TH && tj < TVS_TW && i < n0 && j < n1;
            if (i >= n0) i -= n0;            // periodic wrap (tiles never exceed the image by more than one tile)
            if (j >= n1) j -= n1;
            if (i >= n0) i = 0;
            if (j >= n1) j = 0;
            const int64_t o = i * n1 + j;
            const double2 a = x0[o], gr = g0[o];
            double2 h;
            const double2 y = tv_prox_point(a, gr, tau, h);
            ys[e] = y;
            if (mine) {
                x1[o] = y;
                const double dxx = y.x - a.x, dxy = y.y - a.y, ex = y.x - h.x, ey = y.y - h.y;
                s[0] += dxx * gr.x; s[0] += dxy * gr.y;
                s[1] += dxx * dxx;  s[1] += dxy * dxy;
                s[2] += ex * ex;    s[2] += ey * ey;
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < TVS_TH * TVS_TW; e += TVS_THREADS) {
            const int ti = e / TVS_TW, tj = e - ti * TVS_TW;
            const int64_t i = i0 + ti, j = j0 + tj;
            if (i < n0 && j < n1) {
                const double2 c = ys[ti * (TVS_TW + 1) + tj];
                // neighbours: (i+1, j) and (i, j+1); at the image edge inside a partial tile the wrapped
                // row/column 0 sits at the next tile entry, which was filled with the wrapped index above
                const double2 dn = ys[(ti + 1) * (TVS_TW + 1) + tj];
                const double2 rt = ys[ti * (TVS_TW + 1) + tj + 1];
                const double zi = (dn.x - c.x) + (rt.y - c.y);
                const int64_t o = i * n1 + j;
                double ri, fi;
                loss_elem<LOSS>(zi, b[o], ri, fi);
                r[o] = ri;
                s[3] += fi;
            }
        }
    }
    double* const out[4] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ, scal + FB200_S_F};
    grid_sum<4>(s, red, counter, out);
}

// G = grad(R) fused with the BB reductions, recomputing xhat = x0 - tau*g0 and dx = x1 - x0
template <int BB>
__global__ void __launch_bounds__(TV_THREADS)
tv_grad_bb_fused_kernel(const double* __restrict__ R, int64_t n0, int64_t n1, double2* __restrict__ g,
                        const double2* __restrict__ x0, const double2* __restrict__ g0, const double2* __restrict__ x1,
                        double tau, int strip, double* scal, double* red, unsigned* counter) {
    const int lane   = threadIdx.x & 31;
    const int64_t j  = int64_t(blockIdx.x) * TV_THREADS + threadIdx.x;
    const int64_t i0 = int64_t(blockIdx.y) * strip;
    const bool live  = j < n1;
    const int64_t jc = live ? j : 0;
    const int64_t jl = (jc == 0) ? n1 - 1 : jc - 1;
    double s[3] = {0.0, 0.0, 0.0};
    if (i0 < n0) {
        const int64_t i1 = (i0 + strip < n0) ? i0 + strip : n0;
        double up = R[((i0 == 0) ? n0 - 1 : i0 - 1) * n1 + jc];
        for (int64_t ib = i0; ib < i1; ib += TV_UNROLL) {
            double c[TV_UNROLL], lf[TV_UNROLL];
            double2 p[TV_UNROLL], q[TV_UNROLL], y[TV_UNROLL];
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                const bool ok = i < i1;
                c[u] = ok ? R[i * n1 + jc] : 0.0;
                if (BB >= 2) {
                    p[u] = ok ? x0[i * n1 + jc] : make_double2(0.0, 0.0);
                    q[u] = ok ? g0[i * n1 + jc] : make_double2(0.0, 0.0);
                    y[u] = ok ? x1[i * n1 + jc] : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                double left = __shfl_up_sync(0xffffffffu, c[u], 1);
                if ((lane == 0 || jc == 0) && i < i1) left = R[i * n1 + jl];
                lf[u] = left;
            }
#pragma unroll
            for (int u = 0; u < TV_UNROLL; ++u) {
                const int64_t i = ib + u;
                if (i < i1 && live) {
                    const double a = (u == 0) ? up : c[u - 1];
                    double2 gi;
                    gi.x = a - c[u];
                    gi.y = lf[u] - c[u];
                    g[i * n1 + j] = gi;
                    s[2] += gi.x * gi.x;
                    s[2] += gi.y * gi.y;
                    if (BB >= 2) {
                        const double hx = p[u].x - tau * q[u].x, hy = p[u].y - tau * q[u].y;      // xhat, as in the step kernel
                        const double dxx = y[u].x - p[u].x, dxy = y[u].y - p[u].y;                // dx = x1 - x0
                        const double dg0 = gi.x + (hx - p[u].x) / tau;
                        const double dg1 = gi.y + (hy - p[u].y) / tau;
                        s[0] += dxx * dg0;
                        s[0] += dxy * dg1;
                        s[1] += dg0 * dg0;
                        s[1] += dg1 * dg1;
                    }
                }
            }
            up = c[TV_UNROLL - 1];
        }
    }
    double* const out[3] = {BB >= 2 ? scal + FB200_S_DX_DG : nullptr, BB >= 2 ? scal + FB200_S_DG_SQ : nullptr,
                            scal + FB200_S_G1_SQ};
    grid_sum<3>(s, red, counter, out);
}

// ---------------------------------------------------------------------------------------------------
// Whole TV iteration in ONE pass (non-accelerated modes): forward step, ball projection, divergence,
// loss, AND the speculative gradient with the Barzilai-Borwein sums,
//     xhat = x0 - tau*g0 ; x1 = xhat / max(|xhat|_2, 1) ; r = div(x1) - b ; g1 = grad(r) ; 7 sums
// (reference __init__.py:181-188,248-260,272-274 with tv_denoising.py:26-63,85-96).  Only x0, g0, b are
// read and x1, g1 written: 9U bytes per trial, U = n0*n1*8 (the two-kernel pair above moves 17U).  If the
// line search rejects the trial the kernel simply runs again with the shorter step (as the dense sweep).
//
// Shared-memory stencil with a two-deep halo: for a TH x TW tile the prox point is needed on
// (TH+2) x (TW+2) pixels and the residual on (TH+1) x (TW+1); both live in shared memory.  The tile's own
// x0 / g0 are read again for the BB sums (L1/L2 hits) rather than held in registers, which keeps the
// kernel at 80 registers = 3 blocks (24 warps) per SM: the fp64 sqrt/divide chains of one block then
// overlap the loads of the others.  Every expression is the one of the unfused kernels
// (compiled -fmad=false), so x1 and g1 are bit-identical to them; only the order of the sums differs.
// ---------------------------------------------------------------------------------------------------
constexpr int TVI_TH = 32, TVI_TW = 64, TVI_THREADS = 256;
constexpr int TVI_BLOCKS_PER_SM = 3;
constexpr int TVI_PX = TVI_TH * TVI_TW / TVI_THREADS;               // 8 pixels per thread
constexpr int TVI_YW = TVI_TW + 2, TVI_YH = TVI_TH + 2;             // prox points: rows/cols -1 .. T
constexpr int TVI_RW = TVI_TW + 1, TVI_RH = TVI_TH + 1;             // residuals:   rows/cols -1 .. T-1
constexpr int TVI_SMEM = TVI_YH * TVI_YW * int(sizeof(double2)) + TVI_RH * TVI_RW * int(sizeof(double));

// periodic index: one conditional add/subtract in the common case (|i| within one period), the
// modulo only for images smaller than the tile halo
__device__ __forceinline__ int tv_wrap(int i, int n) {
    if (i < 0) i += n;
    else if (i >= n) i -= n;
    if (i < 0 || i >= n) { i %= n; if (i < 0) i += n; }
    return i;
}

template <int LOSS>
__global__ void __launch_bounds__(TVI_THREADS, TVI_BLOCKS_PER_SM)
tv_iter_kernel(const double2* __restrict__ x0, const double2* __restrict__ g0, double tau, int n0, int n1,
               const double* __restrict__ b, double2* __restrict__ x1, double2* __restrict__ g1, int tiles_x, int ntiles,
               double* scal, double* red, unsigned* counter) {
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) return;
        tau = __ldcg(&scal[FB200_S_TAU]);
    }
    extern __shared__ __align__(16) unsigned char tvi_raw[];
    double2* ys = reinterpret_cast<double2*>(tvi_raw);
    double*  rs = reinterpret_cast<double*>(tvi_raw + TVI_YH * TVI_YW * sizeof(double2));
    // <dx,g0>, <dx,dx>, |x1-xhat|^2, f, <dx,dg>, <dg,dg>, <g1,g1>
    double s[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int tj = threadIdx.x & (TVI_TW - 1), tr = threadIdx.x / TVI_TW;         // own column, first own row
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i0 = (tile / tiles_x) * TVI_TH, j0 = (tile % tiles_x) * TVI_TW;
        const int j = j0 + tj;
        const int jw = tv_wrap(j, n1);
        __syncthreads();                     // the previous tile's reads of ys / rs are done
        // phase 1a: prox point of the tile's own pixels (x1 is final here), step sums
#pragma unroll
        for (int kk = 0; kk < TVI_PX; kk += 4) {
            double2 a[4], gr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t o = int64_t(tv_wrap(i0 + tr + (kk + u) * (TVI_THREADS / TVI_TW), n0)) * n1 + jw;
                a[u]  = x0[o];
                gr[u] = g0[o];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ti = tr + (kk + u) * (TVI_THREADS / TVI_TW);
                const int i = i0 + ti;
                double2 h;
                const double2 y = tv_prox_point(a[u], gr[u], tau, h);
                ys[(ti + 1) * TVI_YW + tj + 1] = y;
                if (i < n0 && j < n1) {
                    x1[int64_t(i) * n1 + j] = y;
                    const double dxx = y.x - a[u].x, dxy = y.y - a[u].y, ex = y.x - h.x, ey = y.y - h.y;
                    s[0] += dxx * gr[u].x; s[0] += dxy * gr[u].y;
                    s[1] += dxx * dxx;     s[1] += dxy * dxy;
                    s[2] += ex * ex;       s[2] += ey * ey;
                }
            }
        }
        // phase 1b: prox point on the halo ring (rows -1 and TH, columns -1 and TW)
        for (int e = threadIdx.x; e < 2 * TVI_YW + 2 * TVI_TH; e += TVI_THREADS) {
            int ti, tc;
            if (e < TVI_YW) { ti = -1; tc = e - 1; }
            else if (e < 2 * TVI_YW) { ti = TVI_TH; tc = e - TVI_YW - 1; }
            else { const int q = e - 2 * TVI_YW; ti = q >> 1; tc = (q & 1) ? TVI_TW : -1; }
            const int64_t o = int64_t(tv_wrap(i0 + ti, n0)) * n1 + tv_wrap(j0 + tc, n1);
            double2 h;
            ys[(ti + 1) * TVI_YW + tc + 1] = tv_prox_point(x0[o], g0[o], tau, h);
        }
        __syncthreads();
        // phase 2: residual r = div(x1) - b on rows/cols -1 .. T-1; f over the tile's own pixels
        for (int e = threadIdx.x; e < TVI_RH * TVI_RW; e += TVI_THREADS) {
            const int ri = e / TVI_RW, rj = e - ri * TVI_RW;          // ti = ri - 1, tj = rj - 1
            const double2 c  = ys[ri * TVI_YW + rj];
            const double2 dn = ys[(ri + 1) * TVI_YW + rj];
            const double2 rt = ys[ri * TVI_YW + rj + 1];
            const double zi = (dn.x - c.x) + (rt.y - c.y);
            const int i = i0 + ri - 1, jj = j0 + rj - 1;
            double rv, fv;
            loss_elem<LOSS>(zi, b[int64_t(tv_wrap(i, n0)) * n1 + tv_wrap(jj, n1)], rv, fv);
            rs[e] = rv;
            if (ri >= 1 && rj >= 1 && i < n0 && jj < n1) s[3] += fv;
        }
        __syncthreads();
        // phase 3: g1 = grad(r) on the tile's own pixels, BB sums
#pragma unroll
        for (int kk = 0; kk < TVI_PX; kk += 4) {
            double2 a[4], gr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t o = int64_t(tv_wrap(i0 + tr + (kk + u) * (TVI_THREADS / TVI_TW), n0)) * n1 + jw;
                a[u]  = x0[o];
                gr[u] = g0[o];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ti = tr + (kk + u) * (TVI_THREADS / TVI_TW);
                const int i = i0 + ti;
                if (i < n0 && j < n1) {
                    const double c = rs[(ti + 1) * TVI_RW + tj + 1];
                    double2 gi;
                    gi.x = rs[ti * TVI_RW + tj + 1] - c;
                    gi.y = rs[(ti + 1) * TVI_RW + tj] - c;
                    g1[int64_t(i) * n1 + j] = gi;
                    const double2 y = ys[(ti + 1) * TVI_YW + tj + 1];
                    const double hx = a[u].x - tau * gr[u].x, hy = a[u].y - tau * gr[u].y;
                    const double dxx = y.x - a[u].x, dxy = y.y - a[u].y;
                    const double dg0 = gi.x + (hx - a[u].x) / tau;
                    const double dg1 = gi.y + (hy - a[u].y) / tau;
                    s[4] += dxx * dg0;   s[4] += dxy * dg1;
                    s[5] += dg0 * dg0;   s[5] += dg1 * dg1;
                    s[6] += gi.x * gi.x; s[6] += gi.y * gi.y;
                }
            }
        }
    }
    double* const out[7] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ, scal + FB200_S_F,
                            scal + FB200_S_DX_DG, scal + FB200_S_DG_SQ, scal + FB200_S_G1_SQ};
    grid_sum<7>(s, red, counter, out);
}

// ---------------------------------------------------------------------------------------------------
// The same whole iteration WITHOUT shared memory or block barriers: a warp owns 30 adjacent image columns
// (lanes 1..30; lanes 0 and 31 carry the halo columns j0-1 and j0+30) and marches down a strip of rows.  The
// prox point of row i+1 is computed while row i is finished, the vertical neighbours live in registers, the
// horizontal ones come from the adjacent lane by shuffle; loads of the next TVM_UNROLL rows are issued before
// the current rows are computed.  No barrier ever separates the fp64 sqrt/divide chains from the loads, which
// is what bounded the tiled kernel above (3 phases per tile at 24 warps/SM).  Same expressions, same bits.
// ---------------------------------------------------------------------------------------------------
constexpr int TVM_THREADS = 128, TVM_COLS = 30;

// v / tau with the reciprocal rt = 1 / tau formed once per thread: product, then two residual corrections
// (Markstein: a faithful quotient corrected with the correctly rounded reciprocal is the correctly rounded
// quotient).  Only the BB sums consume this quotient.
template <bool FD>
__device__ __forceinline__ double tv_div_tau(double v, double tau, double rt) {
    if (!FD) return v / tau;
    const double q0 = v * rt;
    const double q1 = fma(fma(-tau, q0, v), rt, q0);
    return fma(fma(-tau, q1, v), rt, q1);
}

template <int LOSS, int TVM_UNROLL, int MINB, bool FD>
__global__ void __launch_bounds__(TVM_THREADS, MINB)
tv_iter_march_kernel(const double2* __restrict__ x0, const double2* __restrict__ g0, double tau, int n0, int n1,
                     const double* __restrict__ b, double2* __restrict__ x1, double2* __restrict__ g1, int warps_x, int strip,
                     double* scal, double* red, unsigned* counter) {
    const int lane = threadIdx.x & 31;
    const int wid  = blockIdx.x * (TVM_THREADS / 32) + (threadIdx.x >> 5);
    const int wx = wid % warps_x, wy = wid / warps_x;
    const int j  = wx * TVM_COLS - 1 + lane;                  // unwrapped column of this lane
    const int jw = tv_wrap(j, n1);
    const int i0 = wy * strip;
    const int i1 = min(n0, i0 + strip);
    const bool out_lane = lane >= 1 && lane <= TVM_COLS && j < n1;
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) return;
        tau = __ldcg(&scal[FB200_S_TAU]);
    }
    const double rtau = 1.0 / tau;
    double s[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i0 < n0) {
        auto point = [&](int i, double2& a, double2& gr) {    // x0, g0 at (wrapped) row i of this lane's column
            const int64_t o = int64_t(tv_wrap(i, n0)) * n1 + jw;
            a = x0[o];
            gr = g0[o];
        };
        auto resid = [&](double2 c, double2 dn, int i) {      // r = div(y)(i, j) - b(i, j)
            const double rty = __shfl_down_sync(0xffffffffu, c.y, 1);
            const double zi = (dn.x - c.x) + (rty - c.y);
            double rv, fv;
            loss_elem<LOSS>(zi, b[int64_t(tv_wrap(i, n0)) * n1 + jw], rv, fv);
            return make_double2(rv, fv);
        };
        double2 a_c, g_c, a_t, g_t, h;
        point(i0 - 1, a_t, g_t);
        const double2 y_m = tv_prox_point(a_t, g_t, tau, h);       // row i0-1
        point(i0, a_c, g_c);
        double2 y_c = tv_prox_point(a_c, g_c, tau, h);             // row i0
        double r_up = resid(y_m, y_c, i0 - 1).x;                   // r(i0-1, j)
        for (int ib = i0; ib < i1; ib += TVM_UNROLL) {
            double2 an[TVM_UNROLL], gn[TVM_UNROLL];
            double bb[TVM_UNROLL];
#pragma unroll
            for (int u = 0; u < TVM_UNROLL; ++u) {                 // rows ib+1 .. ib+UNROLL and b of rows ib .. ib+UNROLL-1
                const int i = ib + u;
                if (i < i1) {
                    point(i + 1, an[u], gn[u]);
                    bb[u] = b[int64_t(i) * n1 + jw];
                } else {
                    an[u] = gn[u] = make_double2(0.0, 0.0);
                    bb[u] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < TVM_UNROLL; ++u) {
                const int i = ib + u;
                if (i >= i1) break;                                // warp-uniform
                double2 hn;
                const double2 y_n = tv_prox_point(an[u], gn[u], tau, hn);        // row i+1
                const double rty = __shfl_down_sync(0xffffffffu, y_c.y, 1);
                const double zi = (y_n.x - y_c.x) + (rty - y_c.y);
                double rc, fc;
                loss_elem<LOSS>(zi, bb[u], rc, fc);
                const double r_left = __shfl_up_sync(0xffffffffu, rc, 1);
                if (out_lane) {
                    const int64_t o = int64_t(i) * n1 + j;
                    double2 gi;
                    gi.x = r_up - rc;
                    gi.y = r_left - rc;
                    x1[o] = y_c;
                    g1[o] = gi;
                    const double hx = a_c.x - tau * g_c.x, hy = a_c.y - tau * g_c.y;
                    const double dxx = y_c.x - a_c.x, dxy = y_c.y - a_c.y, ex = y_c.x - hx, ey = y_c.y - hy;
                    s[0] += dxx * g_c.x; s[0] += dxy * g_c.y;
                    s[1] += dxx * dxx;   s[1] += dxy * dxy;
                    s[2] += ex * ex;     s[2] += ey * ey;
                    s[3] += fc;
                    const double dg0 = gi.x + tv_div_tau<FD>(hx - a_c.x, tau, rtau);
                    const double dg1 = gi.y + tv_div_tau<FD>(hy - a_c.y, tau, rtau);
                    s[4] += dxx * dg0;   s[4] += dxy * dg1;
                    s[5] += dg0 * dg0;   s[5] += dg1 * dg1;
                    s[6] += gi.x * gi.x; s[6] += gi.y * gi.y;
                }
                r_up = rc;
                y_c = y_n;
                a_c = an[u];
                g_c = gn[u];
            }
        }
    }
    double* const out[7] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ, scal + FB200_S_F,
                            scal + FB200_S_DX_DG, scal + FB200_S_DG_SQ, scal + FB200_S_G1_SQ};
    grid_sum<7>(s, red, counter, out);
}

// ---------------------------------------------------------------------------------------------------
// The marching iteration fed by bulk copies (large images).  tv_iter_march_kernel issues its own global loads, four
// rows ahead, from the same warps that run the fp64 sqrt / divide chains: 16 warps per SM at 124 registers, 57-75 % of
// the HBM rate (long_scoreboard + fixed-latency stalls, profiles/r01_tv_iter_march_full_metrics.csv).  Here a producer
// warp streams the rows of x0, g0 and b of a (strip x 30*CW columns) tile into a shared-memory ring with cp.async.bulk
// (full / empty mbarriers, ~170 KB in flight per SM at zero register cost) and CW compute warps march down the tile
// reading shared memory: same expressions in the same order (tv_prox_point, loss_elem, the Markstein quotient), so
// x1 and g1 are the same bits.  Periodic wrap: the row index wraps in the address; a tile at the left / right image
// edge fetches its halo column with a separate 16-byte copy.  Needs even n1 (16-byte aligned rows of b).
//   ring slot = one image row of the tile:  X[WC] g0-pairs | G[WC] | B[WC + 2],  WC = 30*CW + 2,
//   column c of the image lives at X[c - jL], B[c - jL + 1] with jL = j0 - 1 (the left halo column).
// ---------------------------------------------------------------------------------------------------
template <int CW, int TVT_RING>
struct TvtShape {
    static constexpr int WC = TVM_COLS * CW + 2;
    static constexpr int SLOT = (2 * WC * 16 + (WC + 2) * 8 + 127) / 128 * 128;
    static constexpr int SMEM = TVT_RING * SLOT + 2 * TVT_RING * 8;
    static constexpr int THREADS = (CW + 1) * 32;
};

template <int LOSS, int CW, int U, int TVT_RING>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
tv_iter_tma_kernel(const double2* __restrict__ x0, const double2* __restrict__ g0, double tau, int n0, int n1,
                   const double* __restrict__ b, double2* __restrict__ x1, double2* __restrict__ g1, int tiles_x, int tile_cols,
                   int strip, double* scal, double* red, unsigned* counter) {
    using Sh = TvtShape<CW, TVT_RING>;
    extern __shared__ __align__(128) unsigned char tvt_smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(tvt_smem + TVT_RING * Sh::SLOT);
    uint64_t* empty = full + TVT_RING;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (isnan(tau)) {                                       // speculative trial: see fb200_trial_decide
        if (__ldcg(&scal[FB200_S_SKIP]) != 0.0) return;
        tau = __ldcg(&scal[FB200_S_TAU]);
    }
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int j0 = tx * tile_cols, jL = j0 - 1;             // tile_cols: even, <= 30 * CW; equal tiles, the last one maybe narrower
    const int j_end = min(n1, j0 + tile_cols);              // one past the tile's last column
    const int i0 = ty * strip;
    const int i1 = min(n0, i0 + strip);
    const int nrows = i1 - i0 + 2;                          // image rows i0-1 .. i1 (wrapped)
    if (threadIdx.x == 0) {
        for (int k = 0; k < TVT_RING; ++k) {
            mbar_init(&full[k], 1);
            mbar_init(&empty[k], CW);
        }
        mbar_fence_init();
    }
    __syncthreads();
    double s[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (warp == CW) {
        // ===================================== producer =====================================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const int jR = j_end;                                           // right halo column (unwrapped; n1 at the image edge)
            const int cbeg = jL < 0 ? 0 : jL, cend = jR < n1 ? jR : n1 - 1; // contiguous image columns of the tile
            const uint32_t main_xg = uint32_t(cend - cbeg + 1) * 16u;
            const int bbeg = jL < 0 ? 0 : jL - 1;                           // even (j0 is even)
            int bend = cend;
            if (((bend - bbeg + 1) & 1) && bend + 1 < n1) ++bend;           // even element count (n1 is even)
            const uint32_t main_b = uint32_t(bend - bbeg + 1) * 8u;
            const uint32_t total = 2u * main_xg + main_b + (jL < 0 ? 2u * 16u + 16u : 0u) + (jR >= n1 ? 2u * 16u : 0u);     // jR == n1: the tile touches the right image edge
            for (int k = 0; k < nrows; ++k) {
                const int slot = k % TVT_RING;
                mbar_wait(&empty[slot], ((k / TVT_RING) & 1) ^ 1u);
                unsigned char* base = tvt_smem + size_t(slot) * Sh::SLOT;
                double2* X = reinterpret_cast<double2*>(base);
                double2* G = X + Sh::WC;
                double*  B = reinterpret_cast<double*>(G + Sh::WC);
                const int64_t row = int64_t(tv_wrap(i0 - 1 + k, n0)) * n1;
                mbar_expect_tx(&full[slot], total);
                bulk_load(X + (cbeg - jL), x0 + row + cbeg, main_xg, &full[slot], pol);
                bulk_load(G + (cbeg - jL), g0 + row + cbeg, main_xg, &full[slot], pol);
                bulk_load(B + (bbeg - jL + 1), b + row + bbeg, main_b, &full[slot], pol);
                if (jL < 0) {                                               // left image edge: halo column n1 - 1
                    bulk_load(X, x0 + row + n1 - 1, 16u, &full[slot], pol);
                    bulk_load(G, g0 + row + n1 - 1, 16u, &full[slot], pol);
                    bulk_load(B, b + row + n1 - 2, 16u, &full[slot], pol);  // B[1] = b(., n1 - 1)
                }
                if (jR >= n1) {                                             // right image edge: column 0 follows column n1 - 1
                    bulk_load(X + (n1 - jL), x0 + row, 16u, &full[slot], pol);
                    bulk_load(G + (n1 - jL), g0 + row, 16u, &full[slot], pol);
                }
            }
        }
    } else {
        // ===================================== compute warps: the marching scheme from shared memory ==========
        const int j = j0 + TVM_COLS * warp - 1 + lane;      // unwrapped column of this lane
        const int cx = TVM_COLS * warp + lane;              // its offset in a ring row
        const bool out_lane = lane >= 1 && lane <= TVM_COLS && j < j_end;
        const bool idle_warp = j0 + TVM_COLS * warp >= j_end;      // a warp beyond the tile's columns only keeps the ring turning
        const double rtau = 1.0 / tau;
        auto slot_ptr = [&](int k) { return tvt_smem + size_t(k % TVT_RING) * Sh::SLOT; };
        auto wait_row = [&](int k) { mbar_wait(&full[k % TVT_RING], (k / TVT_RING) & 1); };
        auto free_row = [&](int k) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[k % TVT_RING]);
        };
        auto load_xg = [&](int k, double2& a, double2& gr) {
            const double2* X = reinterpret_cast<const double2*>(slot_ptr(k));
            a = X[cx];
            gr = X[Sh::WC + cx];
        };
        auto load_b = [&](int k) { return reinterpret_cast<const double*>(slot_ptr(k) + 2 * Sh::WC * 16)[cx + 1]; };
        if (idle_warp) {
            for (int k = 0; k < nrows; ++k) {
                wait_row(k);
                free_row(k);
            }
        } else {
        double2 a_c, g_c, a_t, g_t, h;
        wait_row(0);
        load_xg(0, a_t, g_t);
        const double2 y_m = tv_prox_point(a_t, g_t, tau, h);       // row i0-1
        wait_row(1);
        load_xg(1, a_c, g_c);
        double2 y_c = tv_prox_point(a_c, g_c, tau, h);             // row i0
        double r_up;
        {
            const double rty = __shfl_down_sync(0xffffffffu, y_m.y, 1);
            const double zi = (y_c.x - y_m.x) + (rty - y_m.y);
            double fv;
            loss_elem<LOSS>(zi, load_b(0), r_up, fv);              // r(i0-1, j)
        }
        free_row(0);
        for (int ib = i0; ib < i1; ib += U) {
            double2 an[U], gn[U];
            double bb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                          // rows ib+1 .. ib+U and b of rows ib .. ib+U-1
                const int i = ib + u;
                if (i < i1) {
                    const int k = i - i0 + 2;                      // ring index of image row i + 1
                    wait_row(k);
                    load_xg(k, an[u], gn[u]);
                    bb[u] = load_b(k - 1);
                } else {
                    an[u] = gn[u] = make_double2(0.0, 0.0);
                    bb[u] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)                            // rows ib .. ib+U-1 are in registers now
                if (ib + u < i1) free_row(ib + u - i0 + 1);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = ib + u;
                if (i >= i1) break;                                // warp-uniform
                double2 hn;
                const double2 y_n = tv_prox_point(an[u], gn[u], tau, hn);        // row i+1
                const double rty = __shfl_down_sync(0xffffffffu, y_c.y, 1);
                const double zi = (y_n.x - y_c.x) + (rty - y_c.y);
                double rc, fc;
                loss_elem<LOSS>(zi, bb[u], rc, fc);
                const double r_left = __shfl_up_sync(0xffffffffu, rc, 1);
                if (out_lane) {
                    const int64_t o = int64_t(i) * n1 + j;
                    double2 gi;
                    gi.x = r_up - rc;
                    gi.y = r_left - rc;
                    x1[o] = y_c;
                    g1[o] = gi;
                    const double hx = a_c.x - tau * g_c.x, hy = a_c.y - tau * g_c.y;
                    const double dxx = y_c.x - a_c.x, dxy = y_c.y - a_c.y, ex = y_c.x - hx, ey = y_c.y - hy;
                    s[0] += dxx * g_c.x; s[0] += dxy * g_c.y;
                    s[1] += dxx * dxx;   s[1] += dxy * dxy;
                    s[2] += ex * ex;     s[2] += ey * ey;
                    s[3] += fc;
                    const double dg0 = gi.x + tv_div_tau<true>(hx - a_c.x, tau, rtau);
                    const double dg1 = gi.y + tv_div_tau<true>(hy - a_c.y, tau, rtau);
                    s[4] += dxx * dg0;   s[4] += dxy * dg1;
                    s[5] += dg0 * dg0;   s[5] += dg1 * dg1;
                    s[6] += gi.x * gi.x; s[6] += gi.y * gi.y;
                }
                r_up = rc;
                y_c = y_n;
                a_c = an[u];
                g_c = gn[u];
            }
        }
        free_row(nrows - 1);                                       // the last row's slot (its b is never read)
        }
    }
    double* const out[7] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ, scal + FB200_S_F,
                            scal + FB200_S_DX_DG, scal + FB200_S_DG_SQ, scal + FB200_S_G1_SQ};
    grid_sum<7>(s, red, counter, out);
}

// ---------------------------------------------------------------------------------------------------
// FISTA (accelerated) iteration, same marching scheme: forward step and ball projection give the prox point
// x_accel1 (reference __init__.py:181-186), its image z_accel1 = div(x_accel1) and f there feed the line search
// (:187-217); the extrapolation x1 = x_accel1 + c (x_accel1 - x_accel0), z1 = z_accel1 + c (z_accel1 - z_accel0)
// (:242-243, by linearity no second div), f(z1) (:245), the gradient grad(gradf(z1)) (:248) and the BB / residual
// sums (:254-274) follow in the same pass.  The weight c = (alpha0 - 1) / alpha1 depends on the restart test
// <x0 - x_accel1, x_accel1 - x_accel0> > 1e-30 (:231) of THIS trial, which has only two outcomes: the host passes
// the no-restart weight and repeats the launch with c = 0 in the (rare) iterations that restart.
// Reads x0, g0, x_accel0 (2U each), b, z_accel0 (U each); writes x_accel1, x1, g1 (2U each), z_accel1 (U): 15U.
// ---------------------------------------------------------------------------------------------------
template <int LOSS, int TVM_UNROLL, int MINB, bool FD>
__global__ void __launch_bounds__(TVM_THREADS, MINB)
tv_fista_march_kernel(const double2* __restrict__ x0, const double2* __restrict__ g0, double tau, double c, int n0, int n1,
                      const double* __restrict__ b, const double2* __restrict__ xa0, const double* __restrict__ za0,
                      double2* __restrict__ xa1, double* __restrict__ za1, double2* __restrict__ x1,
                      double2* __restrict__ g1, int warps_x, int strip, double* scal, double* red, unsigned* counter) {
    const int lane = threadIdx.x & 31;
    const int wid  = blockIdx.x * (TVM_THREADS / 32) + (threadIdx.x >> 5);
    const int wx = wid % warps_x, wy = wid / warps_x;
    const int j  = wx * TVM_COLS - 1 + lane;                  // unwrapped column of this lane
    const int jw = tv_wrap(j, n1);
    const int i0 = wy * strip;
    const int i1 = min(n0, i0 + strip);
    const bool out_lane = lane >= 1 && lane <= TVM_COLS && j < n1;
    const double rtau = 1.0 / tau;
    double s[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i0 < n0) {
        auto point = [&](int i, double2& a, double2& gr) {    // x0, g0 at (wrapped) row i of this lane's column
            const int64_t o = int64_t(tv_wrap(i, n0)) * n1 + jw;
            a = x0[o];
            gr = g0[o];
        };
        double2 a_c, g_c, a_t, g_t, h;
        point(i0 - 1, a_t, g_t);
        const double2 y_m = tv_prox_point(a_t, g_t, tau, h);       // row i0-1
        point(i0, a_c, g_c);
        double2 y_c = tv_prox_point(a_c, g_c, tau, h);             // row i0
        double r_up;                                               // gradf(z1)(i0-1, j) at the extrapolated z
        {
            const int64_t o = int64_t(tv_wrap(i0 - 1, n0)) * n1 + jw;
            const double rty = __shfl_down_sync(0xffffffffu, y_m.y, 1);
            const double zp = (y_c.x - y_m.x) + (rty - y_m.y);
            const double ze = zp + c * (zp - za0[o]);
            double fv;
            loss_elem<LOSS>(ze, b[o], r_up, fv);
        }
        for (int ib = i0; ib < i1; ib += TVM_UNROLL) {
            double2 an[TVM_UNROLL], gn[TVM_UNROLL], qa[TVM_UNROLL];
            double bb[TVM_UNROLL], zq[TVM_UNROLL];
#pragma unroll
            for (int u = 0; u < TVM_UNROLL; ++u) {                 // x0, g0 of rows ib+1 ..; b, x_accel0, z_accel0 of rows ib ..
                const int i = ib + u;
                if (i < i1) {
                    point(i + 1, an[u], gn[u]);
                    const int64_t o = int64_t(i) * n1 + jw;
                    bb[u] = b[o];
                    zq[u] = za0[o];
                    qa[u] = xa0[o];
                } else {
                    an[u] = gn[u] = qa[u] = make_double2(0.0, 0.0);
                    bb[u] = zq[u] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < TVM_UNROLL; ++u) {
                const int i = ib + u;
                if (i >= i1) break;                                // warp-uniform
                double2 hn;
                const double2 y_n = tv_prox_point(an[u], gn[u], tau, hn);        // row i+1
                const double rty = __shfl_down_sync(0xffffffffu, y_c.y, 1);
                const double zp = (y_n.x - y_c.x) + (rty - y_c.y);               // z_accel1 = div(x_accel1)
                const double ze = zp + c * (zp - zq[u]);                         // extrapolated z1
                double rp, fp, rc, fc;
                loss_elem<LOSS>(zp, bb[u], rp, fp);
                loss_elem<LOSS>(ze, bb[u], rc, fc);
                const double r_left = __shfl_up_sync(0xffffffffu, rc, 1);
                if (out_lane) {
                    const int64_t o = int64_t(i) * n1 + j;
                    double2 gi, xe;
                    gi.x = r_up - rc;
                    gi.y = r_left - rc;
                    xe.x = y_c.x + c * (y_c.x - qa[u].x);
                    xe.y = y_c.y + c * (y_c.y - qa[u].y);
                    xa1[o] = y_c;
                    za1[o] = zp;
                    x1[o] = xe;
                    g1[o] = gi;
                    const double hx = a_c.x - tau * g_c.x, hy = a_c.y - tau * g_c.y;
                    const double dxx = y_c.x - a_c.x, dxy = y_c.y - a_c.y, ex = xe.x - hx, ey = xe.y - hy;
                    s[0] += dxx * g_c.x; s[0] += dxy * g_c.y;
                    s[1] += dxx * dxx;   s[1] += dxy * dxy;
                    s[2] += ex * ex;     s[2] += ey * ey;
                    s[3] += fp;
                    const double dg0 = gi.x + tv_div_tau<FD>(hx - a_c.x, tau, rtau);
                    const double dg1 = gi.y + tv_div_tau<FD>(hy - a_c.y, tau, rtau);
                    s[4] += dxx * dg0;   s[4] += dxy * dg1;
                    s[5] += dg0 * dg0;   s[5] += dg1 * dg1;
                    s[6] += gi.x * gi.x; s[6] += gi.y * gi.y;
                    s[7] += (a_c.x - y_c.x) * (y_c.x - qa[u].x); s[7] += (a_c.y - y_c.y) * (y_c.y - qa[u].y);
                    s[8] += fc;
                }
                r_up = rc;
                y_c = y_n;
                a_c = an[u];
                g_c = gn[u];
            }
        }
    }
    double* const out[9] = {scal + FB200_S_DX_G0, scal + FB200_S_DX_SQ, scal + FB200_S_XMXH_SQ, scal + FB200_S_F,
                            scal + FB200_S_DX_DG, scal + FB200_S_DG_SQ, scal + FB200_S_G1_SQ, scal + FB200_S_RESTART,
                            scal + FB200_S_AUX3};
    grid_sum<9>(s, red, counter, out);
}

static int tv_grid(int64_t n0, int64_t n1, dim3* grid, int* strip) {
    const int64_t gx = (n1 + TV_THREADS - 1) / TV_THREADS;
    int64_t st = TV_STRIP;
    while (gx * ((n0 + st - 1) / st) > MAX_RED_BLOCKS) st *= 2;
    const int64_t gy = (n0 + st - 1) / st;
    if (gx > MAX_RED_BLOCKS || gy > 65535) {
        set_error("tv: image %lld x %lld too large for the stencil grid", (long long)n0, (long long)n1);
        return 1;
    }
    *grid  = dim3(unsigned(gx), unsigned(gy));
    *strip = int(st);
    return 0;
}

// Strip length of the marching kernels: a warp owns TVM_COLS columns x `strip` rows.  The grid should fill a whole
// number of waves of the resident capacity (4 blocks of 4 warps per SM), ending just BELOW a wave boundary, and strips
// should be long enough to amortise the one recomputed halo row: among 1..16 waves pick the strip in [40, 160] rows
// with the least (waves x (strip + halo)) cost.  FASTA_B200_TV_STRIP overrides (experiments).
static int tvm_plan(int64_t n0, int64_t n1, int64_t* warps_x_out, int64_t* strip_out, int64_t* blocks_out) {
    const int64_t warps_x = (n1 + TVM_COLS - 1) / TVM_COLS;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("FASTA_B200_TV_STRIP"); forced = e ? atoi(e) : 0; }
    int64_t strip = 64;
    if (forced > 0) {
        strip = forced;
    } else {
        const int64_t capacity = int64_t(sm_count()) * 4 * (TVM_THREADS / 32);       // resident warps
        double best = 1e300;
        for (int k = 1; k <= 16; ++k) {
            const int64_t strips = (k * capacity) / warps_x;
            if (strips < 1) continue;
            const int64_t st = (n0 + strips - 1) / strips;
            if (st < 40 || st > 160) continue;
            const double cost = double(k) * (double(st) + 1.5);
            if (cost < best - 1e-9) { best = cost; strip = st; }
        }
        if (best > 1e299) {                       // small image: less than one wave at 40 rows -- shorter strips, more warps
            const int64_t strips = capacity / warps_x > 0 ? capacity / warps_x : 1;
            strip = (n0 + strips - 1) / strips;
            strip = strip < 8 ? 8 : (strip > 64 ? 64 : strip);
        }
    }
    while ((warps_x * ((n0 + strip - 1) / strip) + 3) / 4 > MAX_RED_BLOCKS) strip *= 2;
    const int64_t warps = warps_x * ((n0 + strip - 1) / strip);
    const int64_t blocks = (warps + TVM_THREADS / 32 - 1) / (TVM_THREADS / 32);
    if (blocks > MAX_RED_BLOCKS || warps_x > (1 << 24)) return 1;
    *warps_x_out = warps_x; *strip_out = strip; *blocks_out = blocks;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Rank-generic grad / div / ball projection (reference tv_denoising.py:26-63,89-96 are written for N-d arrays; the
// fused kernels above cover the 2-D images of BASELINE config 4, these cover every other rank through the generic
// back-end).  One thread per element; same expressions and summation order as the numpy loops, so the same bits.
// ---------------------------------------------------------------------------------------------------
constexpr int TVN_MAX_RANK = 6;
struct TvnShape { int64_t n[TVN_MAX_RANK]; int64_t stride[TVN_MAX_RANK]; int rank; int64_t total; };

__global__ void __launch_bounds__(256) tv_grad_nd_kernel(const double* __restrict__ X, TvnShape sh, double* __restrict__ G) {
    const int64_t step = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < sh.total; e += step) {
        const double c = X[e];
        for (int d = 0; d < sh.rank; ++d) {
            const int64_t i = (e / sh.stride[d]) % sh.n[d];
            const int64_t prev = (i == 0) ? e + (sh.n[d] - 1) * sh.stride[d] : e - sh.stride[d];
            G[e * sh.rank + d] = X[prev] - c;                 // np.roll(X, 1, axis=d) - X
        }
    }
}

__global__ void __launch_bounds__(256) tv_div_nd_kernel(const double* __restrict__ Y, TvnShape sh, double* __restrict__ D) {
    const int64_t step = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < sh.total; e += step) {
        double acc = 0.0;
        for (int d = 0; d < sh.rank; ++d) {
            const int64_t i = (e / sh.stride[d]) % sh.n[d];
            const int64_t next = (i == sh.n[d] - 1) ? e - (sh.n[d] - 1) * sh.stride[d] : e + sh.stride[d];
            acc += Y[next * sh.rank + d] - Y[e * sh.rank + d];    // divergence += roll(dX, -1, axis=d) - dX
        }
        D[e] = acc;
    }
}

// Y / max(|Y|_2 along the last axis, 1)   (tv_denoising.py:89-96)
__global__ void __launch_bounds__(256) tv_ball_nd_kernel(const double* __restrict__ Y, int64_t npix, int k, double* __restrict__ out) {
    const int64_t step = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < npix; e += step) {
        double ss = 0.0;
        for (int d = 0; d < k; ++d) ss += Y[e * k + d] * Y[e * k + d];
        const double nrm = fmax(sqrt(ss), 1.0);
        for (int d = 0; d < k; ++d) out[e * k + d] = Y[e * k + d] / nrm;
    }
}

static int tvn_shape(const int64_t* shape, int rank, TvnShape* sh) {
    if (rank < 1 || rank > TVN_MAX_RANK) { set_error("tv (N-d): rank %d not in 1..%d", rank, TVN_MAX_RANK); return 1; }
    sh->rank = rank;
    sh->total = 1;
    for (int d = rank - 1; d >= 0; --d) {
        if (shape[d] < 1) { set_error("tv (N-d): empty axis"); return 1; }
        sh->n[d] = shape[d];
        sh->stride[d] = sh->total;
        sh->total *= shape[d];
    }
    return 0;
}

// Tiles of the bulk-copy-fed kernel: tiles_x x strips CTAs, one CTA per SM, the grid a whole number of waves.
template <int CW>
static int tvt_plan(int64_t n0, int64_t n1, int64_t* tiles_x_out, int64_t* tile_cols_out, int64_t* strip_out) {
    const int64_t tiles_x = (n1 + TVM_COLS * CW - 1) / (TVM_COLS * CW);
    // equal column tiles (even width: rows of b stay 16-byte aligned): at 4096 columns 9 tiles of 456 instead of 8 x 480 + 256
    int64_t tile_cols = ((n1 + tiles_x - 1) / tiles_x + 1) / 2 * 2;
    if (tile_cols > TVM_COLS * CW) tile_cols = TVM_COLS * CW;
    const int64_t nsm = sm_count();
    double best = 1e300;
    int64_t strip = 0;
    for (int k = 1; k <= 8; ++k) {
        const int64_t strips = (k * nsm) / tiles_x;
        if (strips < 1) continue;
        const int64_t st = (n0 + strips - 1) / strips;
        if (st < 32 || st > 512) continue;
        const double cost = double(k) * (double(st) + 4.0);         // halo rows + pipeline fill per tile
        if (cost < best - 1e-9) { best = cost; strip = st; }
    }
    *tiles_x_out = tiles_x;
    *tile_cols_out = tile_cols;
    if (!strip) return 1;
    if (tiles_x * ((n0 + strip - 1) / strip) > MAX_RED_BLOCKS) return 1;
    if (double(n1) < 0.75 * double(tiles_x * TVM_COLS * CW)) return 1;     // mostly idle compute warps: not worth it
    *strip_out = strip;
    return 0;
}

// 0: never, 1: large even-width images (default), 2: whenever the layout allows (tests)
static int tvt_mode() {
    const char* e = getenv("FASTA_B200_TV_TMA");
    if (!e) return 1;
    if (e[0] == '0') return 0;
    return (e[0] == 'f' || e[0] == '2') ? 2 : 1;
}

template <int LOSS, int CW, int U, int RING>
static int tvt_launch_as(const double* x0, const double* g0, double tau, int64_t n0, int64_t n1, const double* b, double* x1, double* g1,
                         double* scal, Workspace& w, cudaStream_t st, int mode, bool* done) {
    using Sh = TvtShape<CW, RING>;
    int64_t tiles_x = 0, tile_cols = 0, strip = 0;
    if (tvt_plan<CW>(n0, n1, &tiles_x, &tile_cols, &strip)) {
        if (mode != 2) return 0;
        strip = n0 < 64 ? n0 : 64;                                 // tests: any image, one strip per 64 rows
        if (tiles_x * ((n0 + strip - 1) / strip) > MAX_RED_BLOCKS) return 0;
    }
    static DeviceOnce once;
    if (once.run([] {
            if (cudaFuncSetAttribute(tv_iter_tma_kernel<LOSS, CW, U, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, Sh::SMEM) != cudaSuccess) {
                set_error("tv_iter_tma: shared-memory attribute: %s", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
            return 0;
        }))
        return 1;
    const int64_t blocks = tiles_x * ((n0 + strip - 1) / strip);
    tv_iter_tma_kernel<LOSS, CW, U, RING><<<unsigned(blocks), Sh::THREADS, Sh::SMEM, st>>>(
        (const double2*)x0, (const double2*)g0, tau, int(n0), int(n1), b, (double2*)x1, (double2*)g1, int(tiles_x), int(tile_cols),
        int(strip), scal, w.red, w.counter);
    *done = true;
    return check_launch("tv_iter_tma");
}

template <int LOSS>
static int tvt_launch(const double* x0, const double* g0, double tau, int64_t n0, int64_t n1, const double* b, double* x1, double* g1,
                      double* scal, Workspace& w, cudaStream_t st, bool* done) {
    *done = false;
    const int mode = tvt_mode();
    if (mode == 0 || (n1 & 1) || n1 < 4 || n0 < 1) return 0;
    if ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(g0) | reinterpret_cast<uintptr_t>(b)) & 15) return 0;
    if (mode == 1 && n0 * n1 < (int64_t(1) << 21)) return 0;       // small images: the register-marching kernel
    // measured at 4096^2 (us per launch; compute warps / rows in flight per warp / ring slots): 12/2/12 239.7, 12/4/12 234.2,
    // 14/2/12 218.5, 16/2/10 213.6, 16/4/10 291.4 (spills), 17/2/10 215.9, 18/2/9 235.0, 20/2/8 262.9 (spills); the
    // register-marching kernel 235.7
    int v = 0;
    if (const char* e = getenv("FASTA_B200_TVT_VARIANT")) v = atoi(e);       // experiments
    if (LOSS == FB200_LOSS_LEAST_SQUARES) {
        switch (v) {
            case 1: return tvt_launch_as<LOSS, 12, 2, 12>(x0, g0, tau, n0, n1, b, x1, g1, scal, w, st, mode, done);
            case 2: return tvt_launch_as<LOSS, 14, 2, 12>(x0, g0, tau, n0, n1, b, x1, g1, scal, w, st, mode, done);
            default: break;
        }
    }
    return tvt_launch_as<LOSS, 16, 2, 10>(x0, g0, tau, n0, n1, b, x1, g1, scal, w, st, mode, done);
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_tv_grad_nd(const double* X, const int64_t* shape, int rank, double* G, void* stream) {
    TvnShape sh;
    if (tvn_shape(shape, rank, &sh)) return 1;
    tv_grad_nd_kernel<<<vec_grid(sh.total, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, sh, G);
    return check_launch("tv_grad_nd");
}

extern "C" int fb200_tv_div_nd(const double* Y, const int64_t* shape, int rank, double* D, void* stream) {
    TvnShape sh;
    if (tvn_shape(shape, rank, &sh)) return 1;
    tv_div_nd_kernel<<<vec_grid(sh.total, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(Y, sh, D);
    return check_launch("tv_div_nd");
}

extern "C" int fb200_tv_ball_nd(const double* Y, int64_t npix, int k, double* out, void* stream) {
    if (npix < 1 || k < 1) { set_error("tv_ball_nd: bad shape"); return 1; }
    tv_ball_nd_kernel<<<vec_grid(npix, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(Y, npix, k, out);
    return check_launch("tv_ball_nd");
}

extern "C" int fb200_tv_div_loss(const double* Y, int64_t n0, int64_t n1, int loss, const double* b, double* z,
                                 double* r, double* scal, void* ws, void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid;
    int strip = TV_STRIP;
    if (n0 < 1 || n1 < 1) { set_error("tv_div_loss: bad shape"); return 1; }
    if (tv_grid(n0, n1, &grid, &strip)) return 1;
    switch (loss) {
        case FB200_LOSS_NONE:
            tv_div_loss_kernel<FB200_LOSS_NONE><<<grid, TV_THREADS, 0, st>>>((const double2*)Y, n0, n1, b, z, r, strip, scal, w.red, w.counter);
            break;
        case FB200_LOSS_LEAST_SQUARES:
            tv_div_loss_kernel<FB200_LOSS_LEAST_SQUARES><<<grid, TV_THREADS, 0, st>>>((const double2*)Y, n0, n1, b, z, r, strip, scal, w.red, w.counter);
            break;
        case FB200_LOSS_LOGISTIC:
            tv_div_loss_kernel<FB200_LOSS_LOGISTIC><<<grid, TV_THREADS, 0, st>>>((const double2*)Y, n0, n1, b, z, r, strip, scal, w.red, w.counter);
            break;
        default: set_error("unknown loss tag %d", loss); return 1;
    }
    return check_launch("tv_div_loss");
}

extern "C" int fb200_tv_grad_bb(const double* R, int64_t n0, int64_t n1, double* g, int bb, const double* x0,
                                const double* xhat, const double* dx, double tau, double* scal, void* ws,
                                void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid;
    int strip = TV_STRIP;
    if (n0 < 1 || n1 < 1) { set_error("tv_grad_bb: bad shape"); return 1; }
    if (tv_grid(n0, n1, &grid, &strip)) return 1;
    switch (bb) {
        case 0: tv_grad_bb_kernel<0><<<grid, TV_THREADS, 0, st>>>(R, n0, n1, (double2*)g, (const double2*)x0, (const double2*)xhat, (const double2*)dx, tau, strip, scal, w.red, w.counter); break;
        case 1: tv_grad_bb_kernel<1><<<grid, TV_THREADS, 0, st>>>(R, n0, n1, (double2*)g, (const double2*)x0, (const double2*)xhat, (const double2*)dx, tau, strip, scal, w.red, w.counter); break;
        case 2: tv_grad_bb_kernel<2><<<grid, TV_THREADS, 0, st>>>(R, n0, n1, (double2*)g, (const double2*)x0, (const double2*)xhat, (const double2*)dx, tau, strip, scal, w.red, w.counter); break;
        default: set_error("unknown bb mode %d", bb); return 1;
    }
    return check_launch("tv_grad_bb");
}

extern "C" int fb200_tv_step_div_loss(const double* x0, const double* g0, double tau, int64_t n0, int64_t n1, int loss,
                                      const double* b, double* x1, double* r, double* scal, void* ws, void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n0 < 1 || n1 < 1) { set_error("tv_step_div_loss: bad shape"); return 1; }
    const int64_t tx = (n1 + TVS_TW - 1) / TVS_TW, ty = (n0 + TVS_TH - 1) / TVS_TH;
    if (tx * ty > (int64_t(1) << 30)) { set_error("tv_step_div_loss: image too large"); return 1; }
    const int ntiles = int(tx * ty);
    int grid = sm_count() * 8;
    if (grid > ntiles) grid = ntiles;
    if (grid > MAX_RED_BLOCKS) grid = MAX_RED_BLOCKS;
    switch (loss) {
        case FB200_LOSS_LEAST_SQUARES:
            tv_step_div_loss_kernel<FB200_LOSS_LEAST_SQUARES><<<grid, TVS_THREADS, 0, st>>>((const double2*)x0, (const double2*)g0, tau, n0, n1, b, (double2*)x1, r, int(tx), ntiles, scal, w.red, w.counter);
            break;
        case FB200_LOSS_LOGISTIC:
            tv_step_div_loss_kernel<FB200_LOSS_LOGISTIC><<<grid, TVS_THREADS, 0, st>>>((const double2*)x0, (const double2*)g0, tau, n0, n1, b, (double2*)x1, r, int(tx), ntiles, scal, w.red, w.counter);
            break;
        default: set_error("tv_step_div_loss: unsupported loss tag %d", loss); return 1;
    }
    return check_launch("tv_step_div_loss");
}

extern "C" int fb200_tv_grad_bb_fused(const double* R, int64_t n0, int64_t n1, double* g, int bb, const double* x0,
                                      const double* g0, const double* x1, double tau, double* scal, void* ws,
                                      void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid;
    int strip = TV_STRIP;
    if (n0 < 1 || n1 < 1) { set_error("tv_grad_bb_fused: bad shape"); return 1; }
    if (tv_grid(n0, n1, &grid, &strip)) return 1;
    if (bb >= 2)
        tv_grad_bb_fused_kernel<2><<<grid, TV_THREADS, 0, st>>>(R, n0, n1, (double2*)g, (const double2*)x0, (const double2*)g0, (const double2*)x1, tau, strip, scal, w.red, w.counter);
    else
        tv_grad_bb_fused_kernel<1><<<grid, TV_THREADS, 0, st>>>(R, n0, n1, (double2*)g, (const double2*)x0, (const double2*)g0, (const double2*)x1, tau, strip, scal, w.red, w.counter);
    return check_launch("tv_grad_bb_fused");
}

extern "C" int fb200_tv_iter_fused(const double* x0, const double* g0, double tau, int64_t n0, int64_t n1, int loss,
                                   const double* b, double* x1, double* g1, double* scal, void* ws, void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n0 < 1 || n1 < 1 || n0 > (1 << 30) || n1 > (1 << 30)) { set_error("tv_iter_fused: bad shape"); return 1; }
    static int variant = -1;                // 0: register-marching kernel (default), 1: shared-memory tiles
    if (variant < 0) {
        const char* e = getenv("FASTA_B200_TV_TILED");
        variant = (e && e[0] == '1') ? 1 : 0;
    }
    if (variant == 0) {
        bool done = false;
        if (loss == FB200_LOSS_LEAST_SQUARES) {
            if (tvt_launch<FB200_LOSS_LEAST_SQUARES>(x0, g0, tau, n0, n1, b, x1, g1, scal, w, st, &done)) return 1;
        } else if (loss == FB200_LOSS_LOGISTIC) {
            if (tvt_launch<FB200_LOSS_LOGISTIC>(x0, g0, tau, n0, n1, b, x1, g1, scal, w, st, &done)) return 1;
        }
        if (done) return 0;
        int64_t warps_x, strip, blocks;
        if (tvm_plan(n0, n1, &warps_x, &strip, &blocks)) { set_error("tv_iter_fused: image too large"); return 1; }
        static int mv = -1;                 // experiment knob: unroll depth / occupancy target of the marching kernel
        if (mv < 0) { const char* e = getenv("FASTA_B200_TVM_VARIANT"); mv = e ? atoi(e) : 0; }
#define TVM_LAUNCH(L, U, B) TVM_LAUNCH_FD(L, U, B, false)
#define TVM_LAUNCH_FD(L, U, B, FD) tv_iter_march_kernel<L, U, B, FD><<<unsigned(blocks), TVM_THREADS, 0, st>>>((const double2*)x0, (const double2*)g0, tau, int(n0), int(n1), b, (double2*)x1, (double2*)g1, int(warps_x), int(strip), scal, w.red, w.counter)
#define TVM_PICK(L) switch (mv) { case 1: TVM_LAUNCH(L, 4, 6); break; case 2: TVM_LAUNCH(L, 8, 4); break; case 3: TVM_LAUNCH(L, 2, 8); break; case 4: TVM_LAUNCH(L, 4, 4); break; case 5: TVM_LAUNCH_FD(L, 4, 5, true); break; default: TVM_LAUNCH_FD(L, 4, 4, true); }
        switch (loss) {
            case FB200_LOSS_LEAST_SQUARES: TVM_PICK(FB200_LOSS_LEAST_SQUARES) break;
            case FB200_LOSS_LOGISTIC: TVM_LAUNCH_FD(FB200_LOSS_LOGISTIC, 4, 4, true); break;
            default: set_error("tv_iter_fused: unsupported loss tag %d", loss); return 1;
        }
#undef TVM_PICK
#undef TVM_LAUNCH
#undef TVM_LAUNCH_FD
        return check_launch("tv_iter_march");
    }
    const int64_t tx = (n1 + TVI_TW - 1) / TVI_TW, ty = (n0 + TVI_TH - 1) / TVI_TH;
    if (tx * ty > (int64_t(1) << 30)) { set_error("tv_iter_fused: image too large"); return 1; }
    const int ntiles = int(tx * ty);
    int grid = sm_count() * TVI_BLOCKS_PER_SM;
    if (grid > ntiles) grid = ntiles;
    if (grid > MAX_RED_BLOCKS) grid = MAX_RED_BLOCKS;
    static DeviceOnce attr_once;
    if (attr_once.run([] {
            cudaError_t e1 = cudaFuncSetAttribute(tv_iter_kernel<FB200_LOSS_LEAST_SQUARES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TVI_SMEM);
            cudaError_t e2 = cudaFuncSetAttribute(tv_iter_kernel<FB200_LOSS_LOGISTIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TVI_SMEM);
            if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("tv_iter_fused: smem attribute failed"); cudaGetLastError(); return 1; }
            return 0;
        }))
        return 1;
    switch (loss) {
        case FB200_LOSS_LEAST_SQUARES:
            tv_iter_kernel<FB200_LOSS_LEAST_SQUARES><<<grid, TVI_THREADS, TVI_SMEM, st>>>((const double2*)x0, (const double2*)g0, tau, int(n0), int(n1), b, (double2*)x1, (double2*)g1, int(tx), ntiles, scal, w.red, w.counter);
            break;
        case FB200_LOSS_LOGISTIC:
            tv_iter_kernel<FB200_LOSS_LOGISTIC><<<grid, TVI_THREADS, TVI_SMEM, st>>>((const double2*)x0, (const double2*)g0, tau, int(n0), int(n1), b, (double2*)x1, (double2*)g1, int(tx), ntiles, scal, w.red, w.counter);
            break;
        default: set_error("tv_iter_fused: unsupported loss tag %d", loss); return 1;
    }
    return check_launch("tv_iter_fused");
}

// FISTA trial in one kernel (tv_fista_march_kernel): scal[S_F] = raw f at the prox point's image (line search),
// scal[S_AUX3] = raw f at the extrapolated z, scal[S_RESTART] = the restart dot, the remaining sums as fb200_tv_iter_fused.
extern "C" int fb200_tv_fista_fused(const double* x0, const double* g0, double tau, double c, int64_t n0, int64_t n1, int loss,
                                    const double* b, const double* xa0, const double* za0, double* xa1, double* za1,
                                    double* x1, double* g1, double* scal, void* ws, void* stream) {
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n0 < 1 || n1 < 1 || n0 > (1 << 30) || n1 > (1 << 30)) { set_error("tv_fista_fused: bad shape"); return 1; }
    int64_t warps_x, strip, blocks;
    if (tvm_plan(n0, n1, &warps_x, &strip, &blocks)) { set_error("tv_fista_fused: image too large"); return 1; }
    static int mv = -1;                     // experiment knob: unroll depth / occupancy target
    if (mv < 0) { const char* e = getenv("FASTA_B200_TVF_VARIANT"); mv = e ? atoi(e) : 0; }
#define TVF_LAUNCH(L, U, B) TVF_LAUNCH_FD(L, U, B, false)
#define TVF_LAUNCH_FD(L, U, B, FD) tv_fista_march_kernel<L, U, B, FD><<<unsigned(blocks), TVM_THREADS, 0, st>>>((const double2*)x0, (const double2*)g0, tau, c, int(n0), int(n1), b, (const double2*)xa0, za0, (double2*)xa1, za1, (double2*)x1, (double2*)g1, int(warps_x), int(strip), scal, w.red, w.counter)
    switch (loss) {
        case FB200_LOSS_LEAST_SQUARES:
            switch (mv) { case 1: TVF_LAUNCH_FD(FB200_LOSS_LEAST_SQUARES, 4, 3, true); break; case 2: TVF_LAUNCH(FB200_LOSS_LEAST_SQUARES, 4, 3); break;
                          case 3: TVF_LAUNCH(FB200_LOSS_LEAST_SQUARES, 2, 4); break; default: TVF_LAUNCH_FD(FB200_LOSS_LEAST_SQUARES, 2, 4, true); }
            break;
        case FB200_LOSS_LOGISTIC: TVF_LAUNCH_FD(FB200_LOSS_LOGISTIC, 2, 4, true); break;
        default: set_error("tv_fista_fused: unsupported loss tag %d", loss); return 1;
    }
#undef TVF_LAUNCH
#undef TVF_LAUNCH_FD
    return check_launch("tv_fista_march");
}

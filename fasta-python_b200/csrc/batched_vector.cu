// K14 (vector side): the N x B / M x B kernels of a batch of FBS solves that share A.  Same
// arithmetic as vector_kernels.cu (compiled with -fmad=false: one rounding per numpy operation of
// the reference line), but every column j carries its own step size, threshold and activity flag,
// and every reduction is per column.  Layout: row-major with the batch index fastest, so a warp
// reads 32 consecutive columns of one row (coalesced) and a thread owns ONE column: per-column
// sums need no cross-thread reduction except over the row lanes of a block, and are combined
// across blocks in block order by a tiny finalize kernel (fixed order, no atomics).
#include "common.cuh"

namespace fb200 {

constexpr int BV_THREADS = 256;
constexpr int BV_MAXK    = 5;

struct BvGeom {
    int bc;        // columns handled by a block (<= 256)
    int lanes;     // row lanes per block = 256 / bc
    int rows_per_block;
};

__host__ __device__ inline BvGeom bv_geom(int B) {
    BvGeom g;
    g.bc    = B < BV_THREADS ? B : BV_THREADS;
    g.lanes = BV_THREADS / g.bc;
    g.rows_per_block = 64 * g.lanes;
    return g;
}

// block-level combine of K per-thread column sums over the row lanes, then store to
// part[(blockIdx.x * K + k) * B + col]
template <int K>
__device__ __forceinline__ void bv_store(double (&s)[K], double* part, int B, int col, int lane_row, int lanes, int bc, bool live) {
    __shared__ double sm[BV_MAXK * BV_THREADS];
    const int j = threadIdx.x % bc;
    if (lanes > 1) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (lane_row < lanes) sm[(k * lanes + lane_row) * bc + j] = s[k];
        __syncthreads();
        if (lane_row == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double t = sm[(k * lanes) * bc + j];
                for (int l = 1; l < lanes; ++l) t += sm[(k * lanes + l) * bc + j];
                s[k] = t;
            }
        }
    }
    if (lane_row == 0 && live) {
#pragma unroll
        for (int k = 0; k < K; ++k) part[(size_t(blockIdx.x) * K + k) * B + col] = s[k];
    }
}

// out[k * B + col] = sum over blocks of part[(b * K + k) * B + col] for active columns, in a fixed order:
// 8 row lanes per (k, col) each add the blocks b = lane, lane + 8, ... in order, then the 8 lane sums are
// added in lane order (block = 32 (k,col) pairs x 8 lanes; consecutive threads read consecutive columns)
constexpr int BV_FIN_LANES = 8;
__global__ void __launch_bounds__(256)
bv_finalize_kernel(const double* __restrict__ part, int nblocks, int K, int B, const int* __restrict__ act,
                   double* __restrict__ out) {
    __shared__ double sm[BV_FIN_LANES][32];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + x;
    const bool live = idx < K * B && (!act || act[idx % B]);
    double t = 0.0;
    if (live) {
        const int k = idx / B, col = idx - k * B;
        for (int b = y; b < nblocks; b += BV_FIN_LANES) t += part[(size_t(b) * K + k) * B + col];
    }
    sm[y][x] = t;
    __syncthreads();
    if (y == 0 && live) {
        double r = sm[0][x];
#pragma unroll
        for (int l = 1; l < BV_FIN_LANES; ++l) r += sm[l][x];
        out[idx] = r;
    }
}

// ---- forward step + prox + reductions, per column (reference __init__.py:181-186,200,272-274,285) ----
template <int PROX>
__global__ void __launch_bounds__(BV_THREADS)
bv_fbs_step_kernel(const double* __restrict__ x0, const double* __restrict__ g0, const double* __restrict__ tau,
                   const double* __restrict__ p0v, const double* __restrict__ p1v, const int* __restrict__ act, int64_t n,
                   int B, double* __restrict__ xhat, double* __restrict__ x1, double* __restrict__ dx, double* part) {
    const BvGeom g = bv_geom(B);
    const int j = threadIdx.x % g.bc, lr = threadIdx.x / g.bc;
    const int col = blockIdx.y * g.bc + j;
    const bool live = col < B && lr < g.lanes && act[col];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
        const double t = tau[col], p0 = p0v[col], p1 = p1v ? p1v[col] : 0.0;
        const int64_t r0 = int64_t(blockIdx.x) * g.rows_per_block;
        const int64_t r1 = r0 + g.rows_per_block < n ? r0 + g.rows_per_block : n;
        for (int64_t i = r0 + lr; i < r1; i += g.lanes) {
            const int64_t o = i * B + col;
            const double a = x0[o], gr = g0[o];
            const double h = a - t * gr;
            const double y = prox_elem<PROX>(h, p0, p1);
            const double d = y - a;
            xhat[o] = h;
            x1[o]   = y;
            dx[o]   = d;
            s[0] += d * gr;
            s[1] += d * d;
            const double e = y - h;
            s[2] += e * e;
            s[3] += fabs(y);
        }
    }
    bv_store<4>(s, part, B, col, lr, g.lanes, g.bc, col < B && lr < g.lanes && act[col < B ? col : 0]);
}

// ---- z (sum of S split partials) -> r = gradf(z), f per column (sparse_least_squares.py:41-42) ----
template <int LOSS>
__global__ void __launch_bounds__(BV_THREADS)
bv_loss_kernel(const double* __restrict__ zsrc, int nsplit, int64_t split_stride, const double* __restrict__ b, int b_ld,
               const int* __restrict__ act, int64_t m, int B, double* __restrict__ z, double* __restrict__ r, double* part) {
    const BvGeom g = bv_geom(B);
    const int j = threadIdx.x % g.bc, lr = threadIdx.x / g.bc;
    const int col = blockIdx.y * g.bc + j;
    const bool live = col < B && lr < g.lanes && act[col];
    double s[1] = {0.0};
    if (live) {
        const int64_t r0 = int64_t(blockIdx.x) * g.rows_per_block;
        const int64_t r1 = r0 + g.rows_per_block < m ? r0 + g.rows_per_block : m;
        for (int64_t i = r0 + lr; i < r1; i += g.lanes) {
            const int64_t o = i * B + col;
            double zi = zsrc[o];
            for (int k = 1; k < nsplit; ++k) zi += zsrc[int64_t(k) * split_stride + o];
            z[o] = zi;
            double ri, fi;
            loss_elem<LOSS>(zi, b[b_ld ? i * b_ld + col : i], ri, fi);
            r[o] = ri;
            s[0] += fi;
        }
    }
    bv_store<1>(s, part, B, col, lr, g.lanes, g.bc, col < B && lr < g.lanes && act[col < B ? col : 0]);
}

// ---- g1 (sum of S split partials); dg = g1 + (xhat - x0)/tau; per-column BB sums (__init__.py:254-260,274) ----
__global__ void __launch_bounds__(BV_THREADS)
bv_bb_kernel(const double* __restrict__ gsrc, int nsplit, int64_t split_stride, const double* __restrict__ x0,
             const double* __restrict__ xhat, const double* __restrict__ dx, const double* __restrict__ tau,
             const int* __restrict__ act, int bb, int64_t n, int B, double* __restrict__ g1, double* part) {
    const BvGeom g = bv_geom(B);
    const int j = threadIdx.x % g.bc, lr = threadIdx.x / g.bc;
    const int col = blockIdx.y * g.bc + j;
    const bool live = col < B && lr < g.lanes && act[col];
    double s[3] = {0.0, 0.0, 0.0};
    if (live) {
        const double t = tau ? tau[col] : 1.0;
        const int64_t r0 = int64_t(blockIdx.x) * g.rows_per_block;
        const int64_t r1 = r0 + g.rows_per_block < n ? r0 + g.rows_per_block : n;
        for (int64_t i = r0 + lr; i < r1; i += g.lanes) {
            const int64_t o = i * B + col;
            double gi = gsrc[o];
            for (int k = 1; k < nsplit; ++k) gi += gsrc[int64_t(k) * split_stride + o];
            g1[o] = gi;
            s[2] += gi * gi;
            if (bb >= 2) {
                const double dg = gi + (xhat[o] - x0[o]) / t;
                s[0] += dx[o] * dg;
                s[1] += dg * dg;
            }
        }
    }
    bv_store<3>(s, part, B, col, lr, g.lanes, g.bc, col < B && lr < g.lanes && act[col < B ? col : 0]);
}

// dst[:, j] = src[:, j] for the columns with mask[j] != 0
__global__ void __launch_bounds__(BV_THREADS)
bv_select_kernel(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ mask, int64_t n, int B) {
    const int64_t total = n * B;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t o = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += stride)
        if (mask[o % B]) dst[o] = src[o];
}

// dst[:, dcol[k]] = src[:, scol[k]] for k < npairs (candidate fan-out of the batched line search: a column's
// state is mirrored into spare slots, the winning slot is copied back); thread = (row, pair)
__global__ void __launch_bounds__(BV_THREADS)
bv_copy_cols_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t rows, int ld_dst, int ld_src,
                    const int* __restrict__ scol, const int* __restrict__ dcol, int npairs) {
    const int64_t total = rows * npairs;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t o = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += stride) {
        const int64_t r = o / npairs;
        const int k = int(o - r * npairs);
        dst[r * ld_dst + dcol[k]] = src[r * ld_src + scol[k]];
    }
}

struct BvLaunch {
    dim3 grid;
    int nblocks;
};

static BvLaunch bv_launch(int64_t rows, int B) {
    const BvGeom g = bv_geom(B);
    BvLaunch l;
    l.nblocks = int((rows + g.rows_per_block - 1) / g.rows_per_block);
    l.grid = dim3(unsigned(l.nblocks), unsigned((B + g.bc - 1) / g.bc));
    return l;
}

}  // namespace fb200

using namespace fb200;

extern "C" size_t fb200_batched_workspace_bytes(int64_t M, int64_t N, int64_t B) {
    const int64_t rows = M > N ? M : N;
    const BvGeom g = bv_geom(int(B));
    const int64_t nblocks = (rows + g.rows_per_block - 1) / g.rows_per_block;
    return size_t(nblocks) * BV_MAXK * size_t(B) * sizeof(double) + 1024;
}

#define BV_CHECK_B(B) if ((B) < 1 || (B) > 65535) { set_error("batched: bad batch width"); return 1; }

extern "C" int fb200_batched_fbs_step(const double* x0, const double* g0, const double* tau, int prox, const double* p0,
                                      const double* p1, const int* act, int64_t n, int64_t B, double* xhat, double* x1,
                                      double* dx, double* out, void* ws, void* stream) {
    BV_CHECK_B(B)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const BvLaunch l = bv_launch(n, int(B));
    double* part = static_cast<double*>(ws);
    switch (prox) {
        case FB200_PROX_IDENTITY: bv_fbs_step_kernel<FB200_PROX_IDENTITY><<<l.grid, BV_THREADS, 0, st>>>(x0, g0, tau, p0, p1, act, n, int(B), xhat, x1, dx, part); break;
        case FB200_PROX_SHRINK:   bv_fbs_step_kernel<FB200_PROX_SHRINK><<<l.grid, BV_THREADS, 0, st>>>(x0, g0, tau, p0, p1, act, n, int(B), xhat, x1, dx, part); break;
        case FB200_PROX_NONNEG:   bv_fbs_step_kernel<FB200_PROX_NONNEG><<<l.grid, BV_THREADS, 0, st>>>(x0, g0, tau, p0, p1, act, n, int(B), xhat, x1, dx, part); break;
        case FB200_PROX_BOX:      bv_fbs_step_kernel<FB200_PROX_BOX><<<l.grid, BV_THREADS, 0, st>>>(x0, g0, tau, p0, p1, act, n, int(B), xhat, x1, dx, part); break;
        default: set_error("batched_fbs_step: unsupported prox tag %d", prox); return 1;
    }
    if (check_launch("bv_fbs_step")) return 1;
    bv_finalize_kernel<<<int((4 * B + 31) / 32), 256, 0, st>>>(part, l.nblocks, 4, int(B), act, out);
    return check_launch("bv_finalize");
}

extern "C" int fb200_batched_loss(int loss, const double* zsrc, int nsplit, int64_t split_stride, const double* b,
                                  int64_t b_ld, const int* act, int64_t m, int64_t B, double* z, double* r, double* out,
                                  void* ws, void* stream) {
    BV_CHECK_B(B)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const BvLaunch l = bv_launch(m, int(B));
    double* part = static_cast<double*>(ws);
    if (nsplit < 1) nsplit = 1;
    switch (loss) {
        case FB200_LOSS_LEAST_SQUARES: bv_loss_kernel<FB200_LOSS_LEAST_SQUARES><<<l.grid, BV_THREADS, 0, st>>>(zsrc, nsplit, split_stride, b, int(b_ld), act, m, int(B), z, r, part); break;
        case FB200_LOSS_LOGISTIC:      bv_loss_kernel<FB200_LOSS_LOGISTIC><<<l.grid, BV_THREADS, 0, st>>>(zsrc, nsplit, split_stride, b, int(b_ld), act, m, int(B), z, r, part); break;
        default: set_error("batched_loss: unsupported loss tag %d", loss); return 1;
    }
    if (check_launch("bv_loss")) return 1;
    bv_finalize_kernel<<<int((B + 31) / 32), 256, 0, st>>>(part, l.nblocks, 1, int(B), act, out);
    return check_launch("bv_finalize");
}

extern "C" int fb200_batched_bb(const double* gsrc, int nsplit, int64_t split_stride, const double* x0, const double* xhat,
                                const double* dx, const double* tau, const int* act, int bb, int64_t n, int64_t B,
                                double* g1, double* out, void* ws, void* stream) {
    BV_CHECK_B(B)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const BvLaunch l = bv_launch(n, int(B));
    double* part = static_cast<double*>(ws);
    if (nsplit < 1) nsplit = 1;
    bv_bb_kernel<<<l.grid, BV_THREADS, 0, st>>>(gsrc, nsplit, split_stride, x0, xhat, dx, tau, act, bb, n, int(B), g1, part);
    if (check_launch("bv_bb")) return 1;
    bv_finalize_kernel<<<int((3 * B + 31) / 32), 256, 0, st>>>(part, l.nblocks, 3, int(B), act, out);
    return check_launch("bv_finalize");
}

extern "C" int fb200_batched_select(double* dst, const double* src, const int* mask, int64_t n, int64_t B, void* stream) {
    BV_CHECK_B(B)
    const int64_t total = n * B;
    int grid = int((total + BV_THREADS * 4 - 1) / (BV_THREADS * 4));
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    bv_select_kernel<<<grid, BV_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, mask, n, int(B));
    return check_launch("bv_select");
}

extern "C" int fb200_batched_copy_cols(double* dst, const double* src, int64_t rows, int64_t ld_dst, int64_t ld_src,
                                       const int* scol, const int* dcol, int npairs, void* stream) {
    if (rows < 1 || npairs < 1) return 0;
    const int64_t total = rows * npairs;
    int grid = int((total + BV_THREADS - 1) / BV_THREADS);
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    bv_copy_cols_kernel<<<grid, BV_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, rows, int(ld_dst), int(ld_src), scol, dcol, npairs);
    return check_launch("bv_copy_cols");
}

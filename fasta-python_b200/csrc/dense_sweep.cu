// Single-pass FBS sweep over a dense row-major fp64 matrix:
//
//     z = A x ,  r = gradf(z) ,  f = sum loss(z) ,  g = A^T r        -- ONE read of A from HBM.
//
// The reference does this as two separate numpy contractions per iteration (linalg.py:41 `A @ x`
// at __init__.py:187 and `A.T @ r` at :248) with `f(z)`/`gradf(z)` between them (:188,:248); both
// stream the same A at the same point x1, so the second HBM pass is avoidable if a row of A can
// stay on chip between "dot it with x" and "scale it by r_i and add it to g".  A row (800 KB at
// N=100000) does not fit one SM, but it fits the distributed shared memory of a thread-block
// CLUSTER:
//
//   * a cluster of CS CTAs owns a contiguous range of rows; CTA `rank` owns the column slab
//     [rank*Nc, (rank+1)*Nc) of every row, Nc = ceil(N / CS) <= 6656;
//   * producer warp: one cp.async.bulk (TMA, no tensor map; L2 evict-first) per row slab into a
//     ring of NST stages (<= 53 KB each) guarded by full/empty mbarriers;
//   * A-group (8 warps): holds its x slab in REGISTERS for the whole kernel; per row it dots the
//     slab with x (conflict-free LDS.128 + DFMA), block-reduces, and the CTA's partial is pushed
//     to every CTA of the cluster with st.async (one-way DSMEM store that completes a tx-count on
//     the receiver's mbarrier -- no cluster-wide barrier in the steady state);
//   * B-group (8 warps): holds its g slab accumulators in REGISTERS; per row it waits for the CS
//     partials, adds them in rank order (every CTA gets the bit-identical z_i), evaluates the loss
//     (r_i = gradf(z_i)), and re-reads the slab from shared memory: g += A[i, slab] * r_i;
//   * at the end each cluster writes its g partial; the Barzilai-Borwein epilogue kernel adds the
//     cluster partials in index order (bit-reproducible, no atomics).
//
// Flow control: a sender can run at most 2*NST rows ahead of any receiver's consumption (its A-group
// is bounded by its own stage ring, which its B-group frees only after receiving everybody's
// partials), so NSLOT = 16 >= 2*NST receive slots never alias.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fb200 {

int launch_bb(int bb, const double* gsrc, int nsplit, int64_t ld, int64_t n, double* g, const double* x0,
              const double* xhat, const double* dx, double tau, double* scal, Workspace& w, cudaStream_t st);

constexpr int SW_GROUP   = 256;                 // threads in the A-group and in the B-group
constexpr int SW_THREADS = 2 * SW_GROUP + 32;   // + producer warp
constexpr int SW_NSLOT   = 16;
constexpr int SW_MAXCS   = 16;
constexpr int SW_MAXSTG  = 8;
constexpr int SW_TAIL    = SW_NSLOT * SW_MAXCS * 8 + 2 * 8 * 8 + (2 * SW_MAXSTG + SW_NSLOT) * 8;   // recv + redA + barriers
constexpr int SW_SMEM_MAX = 227 * 1024;

template <int LOSS>
__device__ __forceinline__ void loss_strict(double z, double b, double& r, double& f) {
    if (LOSS == FB200_LOSS_LEAST_SQUARES) {
        r = __dsub_rn(z, b);
        f = __dmul_rn(r, r);
    } else if (LOSS == FB200_LOSS_LOGISTIC) {
        const double ind = (b == 1.0) ? 1.0 : 0.0;
        f = __dsub_rn(log(__dadd_rn(1.0, exp(z))), __dmul_rn(ind, z));
        r = __ddiv_rn(-b, __dadd_rn(1.0, exp(__dmul_rn(b, z))));
    } else {
        r = z;      // LOSS_NONE: "gradient" is z itself (g = A^T A x), f unused
        f = 0.0;
    }
}

template <int LOSS, int CPT>
__global__ void __launch_bounds__(SW_THREADS, 1)
dense_sweep_kernel(const double* __restrict__ A, int64_t lda, int M, int N, int Nc, const double* __restrict__ x,
                   const double* __restrict__ b, double* __restrict__ z, double* __restrict__ r,
                   double* __restrict__ gpart, int64_t ldg, double* __restrict__ fpart, int nstage,
                   const double* __restrict__ za0, double* __restrict__ za1, double cacc, double* __restrict__ fpart2,
                   const double* __restrict__ skip) {
    if (skip && __ldcg(skip) != 0.0) return;       // speculative trial that must not run (fb200_trial_decide): whole grid
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int STAGE_BYTES = CPT * SW_GROUP * 16;
    double*   recv  = reinterpret_cast<double*>(smem + size_t(nstage) * STAGE_BYTES);     // [NSLOT][MAXCS]
    double*   redA  = recv + SW_NSLOT * SW_MAXCS;                                          // [2][8]
    uint64_t* full  = reinterpret_cast<uint64_t*>(redA + 16);
    uint64_t* empty = full + SW_MAXSTG;
    uint64_t* rbar  = empty + SW_MAXSTG;                                                   // [NSLOT]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank(), csize = cluster_nctarank();
    const uint32_t cid = cluster_id_x(), ncl = cluster_count_x();
    const int row_lo = int((int64_t(cid) * M) / ncl), row_hi = int((int64_t(cid + 1) * M) / ncl);
    const int c0 = int(rank) * Nc;                               // first column of this CTA's slab
    int ncols = N - c0;
    ncols = ncols < 0 ? 0 : (ncols > Nc ? Nc : ncols);           // even (N and Nc are even)
    const uint32_t slab_bytes = uint32_t(ncols) * 8u;

    // zero the stage ring once: the tail beyond `ncols` is never written by the bulk copies
    // ... and the receive table: entries of ranks >= csize are never written and must add as zeros
    for (int i = tid; i < (nstage * STAGE_BYTES + SW_NSLOT * SW_MAXCS * 8) / 16; i += SW_THREADS)
        reinterpret_cast<double2*>(smem)[i] = make_double2(0.0, 0.0);
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SW_GROUP / 32);     // one arrival per B-group warp
        }
        for (int s = 0; s < SW_NSLOT; ++s) mbar_init(&rbar[s], 1);
        mbar_fence_init();
    }
    fence_proxy_async();        // generic-proxy zero fill before async-proxy (bulk copy) writes
    __syncthreads();
    cluster_sync_all();         // every CTA's barriers exist before anyone signals them remotely

    if (warp == 2 * SW_GROUP / 32) {
        // ===================================== producer =====================================
        if (lane == 0 && slab_bytes > 0) {
            const uint64_t pol = policy_evict_first();
            int stage = 0;
            uint32_t phase = 0;
            for (int row = row_lo; row < row_hi; ++row) {
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_expect_tx(&full[stage], slab_bytes);
                bulk_load(smem + size_t(stage) * STAGE_BYTES, A + int64_t(row) * lda + c0, slab_bytes, &full[stage], pol);
                if (++stage == nstage) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp < SW_GROUP / 32) {
        // ===================================== A-group: partial z = slab . x ==========================
        double2 xr[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = c0 + 2 * (tid + SW_GROUP * k);
            xr[k] = (col < c0 + ncols) ? *reinterpret_cast<const double2*>(x + col) : make_double2(0.0, 0.0);
        }
        // remote addresses of my column in every peer's receive table, and of their barriers
        uint32_t peer_recv = 0, peer_bar = 0;
        if (warp == 0 && lane < int(csize)) {
            peer_recv = map_to_rank(smem_u32(recv + rank), lane);
            peer_bar  = map_to_rank(smem_u32(rbar), lane);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int row = row_lo; row < row_hi; ++row) {
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            if (slab_bytes > 0) {
                mbar_wait(&full[stage], phase);
                const double2* a = reinterpret_cast<const double2*>(smem + size_t(stage) * STAGE_BYTES) + tid;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const double2 v = a[SW_GROUP * k];
                    if (k & 1) {
                        p2 = fma(v.x, xr[k].x, p2);
                        p3 = fma(v.y, xr[k].y, p3);
                    } else {
                        p0 = fma(v.x, xr[k].x, p0);
                        p1 = fma(v.y, xr[k].y, p1);
                    }
                }
            }
            double p = warp_sum((p0 + p1) + (p2 + p3));
            const int par = (row - row_lo) & 1;
            if (lane == 0) redA[par * 8 + warp] = p;
            asm volatile("bar.sync 1, %0;" ::"n"(SW_GROUP) : "memory");
            if (warp == 0) {
                double v = redA[par * 8 + (lane & 7)];
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // the 8 warp partials, fixed order
                if (lane < int(csize)) {
                    const int slot = (row - row_lo) & (SW_NSLOT - 1);
                    st_async_f64(peer_recv + uint32_t(slot) * SW_MAXCS * 8u, v, peer_bar + uint32_t(slot) * 8u);
                }
            }
            if (++stage == nstage) { stage = 0; phase ^= 1u; }
        }
    } else {
        // ===================================== B-group: z, loss, g += slab * r_i ======================
        const int t = tid - SW_GROUP;
        const int bw = warp - SW_GROUP / 32;
        double2 gr[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) gr[k] = make_double2(0.0, 0.0);
        double facc = 0.0;
        int stage = 0;
        uint32_t phase = 0;
        double bnext = (row_lo < row_hi && b) ? __ldg(b + row_lo) : 0.0;
        // FISTA mode (za0 != nullptr): x is the prox point, z_i its image; the gradient is taken at the
        // extrapolated z_i + c (z_i - za0_i) (reference __init__.py:244-248), a per-row function of z_i
        double qnext = (row_lo < row_hi && za0) ? __ldg(za0 + row_lo) : 0.0;
        double facc2 = 0.0;
        for (int row = row_lo; row < row_hi; ++row) {
            const int it = row - row_lo;
            const int slot = it & (SW_NSLOT - 1);
            const uint32_t rpar = uint32_t(it / SW_NSLOT) & 1u;
            if (t == 0) mbar_expect_tx(&rbar[slot], csize * 8u);
            const double bi = bnext, qi = qnext;
            if (row + 1 < row_hi && b) bnext = __ldg(b + row + 1);
            if (row + 1 < row_hi && za0) qnext = __ldg(za0 + row + 1);
            // st.async data is visible to whoever observes the phase completion of the barrier it signals
            // (a cluster-scope acquire here would make ptxas emit an L1 invalidate, CCTL.IVALL, per row)
            mbar_wait(&rbar[slot], rpar);
            const double2* rv = reinterpret_cast<const double2*>(recv + slot * SW_MAXCS);
            double zi;      // fixed binary tree over ranks: same bits on every CTA (missing ranks hold zeros)
            {
                const double2 e0 = rv[0], e1 = rv[1], e2 = rv[2], e3 = rv[3];
                const double s0 = (e0.x + e0.y) + (e1.x + e1.y), s1 = (e2.x + e2.y) + (e3.x + e3.y);
                const double2 e4 = rv[4], e5 = rv[5], e6 = rv[6], e7 = rv[7];
                const double s2 = (e4.x + e4.y) + (e5.x + e5.y), s3 = (e6.x + e6.y) + (e7.x + e7.y);
                zi = (s0 + s1) + (s2 + s3);
            }
            double ri, fi;
            const double ze = za0 ? __dadd_rn(zi, __dmul_rn(cacc, __dsub_rn(zi, qi))) : zi;   // p + c*(p - za0), as accel_step
            loss_strict<LOSS>(ze, bi, ri, fi);
            if (slab_bytes > 0) {
                mbar_wait(&full[stage], phase);      // already complete (the A-group saw it); orders the TMA data for us
                const double2* a = reinterpret_cast<const double2*>(smem + size_t(stage) * STAGE_BYTES) + t;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const double2 v = a[SW_GROUP * k];
                    gr[k].x = fma(v.x, ri, gr[k].x);
                    gr[k].y = fma(v.y, ri, gr[k].y);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
            if (rank == 0 && t == 0) {
                facc = __dadd_rn(facc, fi);
                z[row] = ze;
                if (LOSS != FB200_LOSS_NONE) r[row] = ri;
                if (za0) {                           // f at the prox point (line search) and the prox image itself
                    double rp, fp;
                    loss_strict<LOSS>(zi, bi, rp, fp);
                    facc2 = __dadd_rn(facc2, fp);
                    za1[row] = zi;
                }
            }
            if (++stage == nstage) { stage = 0; phase ^= 1u; }
        }
        (void)bw;
        double* gp = gpart + int64_t(cid) * ldg;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = c0 + 2 * (t + SW_GROUP * k);
            if (col < c0 + ncols) *reinterpret_cast<double2*>(gp + col) = gr[k];
        }
        if (rank == 0 && t == 0) {
            fpart[cid] = facc;
            if (za0) fpart2[cid] = facc2;
        }
    }
    cluster_sync_all();         // nobody leaves while a peer might still address its shared memory
}

// sum of the per-cluster loss partials in index order -> scal[S_F]
__global__ void sweep_fsum_kernel(const double* __restrict__ fpart, int n, double* out, const double* skip) {
    if (skip && __ldcg(skip) != 0.0) return;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += fpart[i];
        *out = s;
    }
}

struct SweepPlan {
    bool ok;
    int cs, cpt, nc, nstage, ncl, smem;
};

typedef void (*SweepKernel)(const double*, int64_t, int, int, int, const double*, const double*, double*, double*,
                            double*, int64_t, double*, int, const double*, double*, double, double*, const double*);

template <int LOSS>
static SweepKernel pick_kernel(int cpt) {
    switch (cpt) {
        case 4: return dense_sweep_kernel<LOSS, 4>;
        case 7: return dense_sweep_kernel<LOSS, 7>;
        case 10: return dense_sweep_kernel<LOSS, 10>;
        default: return dense_sweep_kernel<LOSS, 13>;
    }
}

static SweepKernel kernel_for(int loss, int cpt) {
    switch (loss) {
        case FB200_LOSS_LEAST_SQUARES: return pick_kernel<FB200_LOSS_LEAST_SQUARES>(cpt);
        case FB200_LOSS_LOGISTIC: return pick_kernel<FB200_LOSS_LOGISTIC>(cpt);
        default: return pick_kernel<FB200_LOSS_NONE>(cpt);
    }
}

static int fill_launch(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, const SweepPlan& p, int ncl, cudaStream_t st) {
    cfg->gridDim          = dim3(unsigned(ncl * p.cs));
    cfg->blockDim         = dim3(SW_THREADS);
    cfg->dynamicSmemBytes = size_t(p.smem);
    cfg->stream           = st;
    attr[0].id                 = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x   = unsigned(p.cs);
    attr[0].val.clusterDim.y   = 1;
    attr[0].val.clusterDim.z   = 1;
    cfg->attrs    = attr;
    cfg->numAttrs = 1;
    return 0;
}

// plan cache keyed by (N, M-independent) -- the occupancy query costs ~10 us, the attribute set more
static SweepPlan make_sweep_plan(int loss, int64_t M, int64_t N) {
    // (per device: the attributes set below and the occupancy answer belong to the current device's context)
    static std::mutex mtx;
    static SweepPlan cache[3][64];
    static int64_t cache_n[3][64];
    static int cache_dev[3][64];
    static int cache_used[3] = {0, 0, 0};
    const int li = loss == FB200_LOSS_LEAST_SQUARES ? 1 : (loss == FB200_LOSS_LOGISTIC ? 2 : 0);
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mtx);
    for (int i = 0; i < cache_used[li]; ++i)
        if (cache_n[li][i] == N && cache_dev[li][i] == dev) return cache[li][i];

    SweepPlan p{};
    p.ok = false;
    const int cpts[4] = {4, 7, 10, 13};
    for (int cs = 1; cs <= SW_MAXCS && !p.ok; cs *= 2) {
        const int64_t nc = round_up((N + cs - 1) / cs, 2);
        if (nc > int64_t(SW_GROUP) * 2 * 13) continue;
        int cpt = 13;
        for (int c : cpts)
            if (nc <= int64_t(SW_GROUP) * 2 * c) { cpt = c; break; }
        const int stage_bytes = cpt * SW_GROUP * 16;
        int nst = (SW_SMEM_MAX - SW_TAIL - 256) / stage_bytes;
        if (nst > SW_MAXSTG) nst = SW_MAXSTG;
        if (nst < 2) continue;
        p.cs = cs; p.cpt = cpt; p.nc = int(nc); p.nstage = nst;
        p.smem = nst * stage_bytes + SW_TAIL;
        SweepKernel k = kernel_for(loss, cpt);
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem) != cudaSuccess) { cudaGetLastError(); continue; }
        if (cs > 8 && cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        fill_launch(&cfg, attr, p, 1, nullptr);
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, k, &cfg) != cudaSuccess || ncl < 1) { cudaGetLastError(); continue; }
        p.ncl = ncl;
        p.ok  = true;
    }
    if (cache_used[li] < 64) {
        cache_n[li][cache_used[li]] = N;
        cache_dev[li][cache_used[li]] = dev;
        cache[li][cache_used[li]++] = p;
    }
    (void)M;
    return p;
}

static bool sweep_eligible(const double* A, int64_t lda, int64_t M, int64_t N) {
    return (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % 2 == 0) && (N % 2 == 0) && lda >= N && M > 0 && N > 0 &&
           M < (int64_t(1) << 31) && N <= int64_t(SW_MAXCS) * SW_GROUP * 2 * 13;
}

int sweep_max_clusters() { return 160; }

// grid variant (dense_gsweep.cu): all SMs, partial dots exchanged through L2 instead of cluster DSMEM
bool gsweep_eligible(const double* A, int64_t lda, int64_t M, int64_t N);
int gsweep_plan(int64_t M, int64_t N, int* plan);
int gsweep_launch(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss, const double* b,
                  double* z, double* r, double* g, int bb, const double* x0, const double* xhat, const double* dx,
                  double tau, double* scal, void* ws, size_t ws_bytes, void* stream, const double* za0,
                  double* za1, double c, int64_t* raw);

// FASTA_B200_SWEEP_KERNEL=cluster selects the cluster / DSMEM kernel of this file (the round-1 default)
static bool use_grid_sweep() {
    const char* e = getenv("FASTA_B200_SWEEP_KERNEL");
    return !(e && e[0] == 'c');
}

}  // namespace fb200

using namespace fb200;

// plan[0..4] = cluster size, clusters co-resident, stages, columns per CTA, double2 per thread
extern "C" int fb200_sweep_plan(int64_t M, int64_t N, int* plan) {
    if (use_grid_sweep()) return gsweep_plan(M, N, plan);      // slabs, bands, stages, columns per CTA, double2 per thread
    SweepPlan p = make_sweep_plan(FB200_LOSS_LEAST_SQUARES, M, N);
    if (!p.ok) return 1;
    plan[0] = p.cs; plan[1] = p.ncl; plan[2] = p.nstage; plan[3] = p.nc; plan[4] = p.cpt;
    return 0;
}

extern "C" int fb200_sweep_supported(const double* A, int64_t lda, int64_t M, int64_t N) {
    if (use_grid_sweep()) {
        int plan[5];
        return (gsweep_eligible(A, lda, M, N) && gsweep_plan(M, N, plan) == 0) ? plan[0] : 0;
    }
    if (!sweep_eligible(A, lda, M, N)) return 0;
    SweepPlan p = make_sweep_plan(FB200_LOSS_LEAST_SQUARES, M, N);
    return p.ok ? p.cs : 0;
}

// za0 != nullptr: FISTA mode (see the kernel); S_F then holds f at the prox point and S_AUX3 f at the extrapolated z
static int sweep_launch(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss, const double* b,
                        double* z, double* r, double* g, int bb, const double* x0, const double* xhat, const double* dx,
                        double tau, double* scal, void* ws, size_t ws_bytes, void* stream, const double* za0,
                        double* za1, double c) {
    if (use_grid_sweep())
        return gsweep_launch(A, lda, M, N, x, loss, b, z, r, g, bb, x0, xhat, dx, tau, scal, ws, ws_bytes, stream, za0, za1, c, nullptr);
    if (!sweep_eligible(A, lda, M, N) || reinterpret_cast<uintptr_t>(x) % 16 != 0) {
        set_error("dense_sweep: matrix not eligible (needs 16-byte aligned base and x, even lda and N, N <= %d)", SW_MAXCS * SW_GROUP * 26);
        return 1;
    }
    if (ws_bytes < fb200_workspace_bytes(M, N)) { set_error("dense_sweep: workspace too small"); return 1; }
    SweepPlan p = make_sweep_plan(loss, M, N);
    if (!p.ok) { set_error("dense_sweep: no feasible cluster configuration for N=%lld", (long long)N); return 1; }
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t ldg = round_up(N, 256);
    int ncl = p.ncl;
    if (const char* e = getenv("FB200_SWEEP_NCL")) { int v = atoi(e); if (v > 0 && v < ncl) ncl = v; }   // experiments only
    if (ncl > M) ncl = int(M);
    const int64_t cap = int64_t(fb200_workspace_bytes(M, N) - DENSE_OFF) / 8 / ldg;
    if (ncl > cap) ncl = int(cap);
    const int fcap = za0 ? FPART_MAX / 2 : FPART_MAX;       // FISTA mode keeps two loss partials per cluster
    if (ncl > fcap) ncl = fcap;
    SweepKernel k = kernel_for(loss, p.cpt);
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    fill_launch(&cfg, attr, p, ncl, st);
    double* fpart = w.fpart;
    double* fpart2 = w.fpart + FPART_MAX / 2;
    const double* skip = isnan(tau) ? scal + FB200_S_SKIP : nullptr;      // speculative trial: see fb200_trial_decide
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, A, lda, int(M), int(N), p.nc, x, b, z, r, w.dense, ldg, fpart, p.nstage, za0, za1,
                                       c, fpart2, skip);
    if (e != cudaSuccess) { set_error("dense_sweep: launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    if (loss != FB200_LOSS_NONE) {
        if (za0) {
            sweep_fsum_kernel<<<1, 32, 0, st>>>(fpart2, ncl, scal + FB200_S_F, skip);       // prox point: the line-search value
            sweep_fsum_kernel<<<1, 32, 0, st>>>(fpart, ncl, scal + FB200_S_AUX3, skip);     // extrapolated point
        } else {
            sweep_fsum_kernel<<<1, 32, 0, st>>>(fpart, ncl, scal + FB200_S_F, skip);
        }
        if (check_launch("sweep_fsum_kernel")) return 1;
    }
    // Barzilai-Borwein epilogue: fixed-order sum of the cluster partials (+ reductions)
    if (g) return launch_bb(bb, w.dense, ncl, ldg, N, g, x0, xhat, dx, tau, scal, w, st);
    return 0;
}

extern "C" int fb200_dense_sweep(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss,
                                 const double* b, double* z, double* r, double* g, int bb, const double* x0,
                                 const double* xhat, const double* dx, double tau, double* scal, void* ws,
                                 size_t ws_bytes, void* stream) {
    return sweep_launch(A, lda, M, N, x, loss, b, z, r, g, bb, x0, xhat, dx, tau, scal, ws, ws_bytes, stream, nullptr, nullptr, 0.0);
}

// ---- row-sharded map: the sweep on this rank's rows + ONE kernel that finishes it across the ranks -------------------
namespace fb200 {
int launch_peer_exchange(const uint64_t* peer_data, const uint64_t* peer_flags, int rank, int P, uint32_t epoch, int64_t n,
                         const double* gsrc, int nsplit, int64_t ld, const double* fpart, const double* fpart2, int with_loss,
                         double* g, int bb, const double* x0, const double* xhat, const double* dx, double tau,
                         const int* decide_i, const double* decide_d, double* host_out, double* scal, Workspace& w, cudaStream_t st);
}

extern "C" int fb200_sweep_exchange_supported(void) { return use_grid_sweep() ? 1 : 0; }

extern "C" int fb200_dense_sweep_exchange(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss,
                                          const double* b, double* z, double* r, const double* za0, double c, double* za1,
                                          const uint64_t* peer_data, const uint64_t* peer_flags, int rank, int P,
                                          uint32_t epoch, double* g, int bb, const double* x0, const double* xhat,
                                          const double* dx, double tau, const int* decide_i, const double* decide_d,
                                          double* host_out, double* scal, void* ws, size_t ws_bytes, void* stream) {
    if (!use_grid_sweep()) { set_error("dense_sweep_exchange: needs the grid sweep kernel"); return 1; }
    if (P < 1 || P > FB200_MAX_PEERS || rank < 0 || rank >= P) { set_error("dense_sweep_exchange: bad rank / world size"); return 1; }
    int64_t raw[2] = {0, 0};
    if (gsweep_launch(A, lda, M, N, x, loss, b, z, r, nullptr, 0, nullptr, nullptr, nullptr, isnan(tau) ? tau : 0.0, scal, ws, ws_bytes,
                      stream, za0, za1, c, raw))
        return 1;
    Workspace w(ws);
    const int with_loss = loss == FB200_LOSS_NONE ? 0 : (za0 ? 2 : 1);
    // FISTA mode: fpart2 holds the loss at the prox point (the line-search value, S_F), fpart at the extrapolated point (S_AUX3)
    const double* f_first = za0 ? w.fpart + FPART_MAX / 2 : w.fpart;
    const double* f_second = za0 ? w.fpart : nullptr;
    return launch_peer_exchange(peer_data, peer_flags, rank, P, epoch, N, w.dense, int(raw[0]), raw[1], f_first, f_second, with_loss,
                                g, bb, x0, xhat, dx, tau, decide_i, decide_d, host_out, scal, w, static_cast<cudaStream_t>(stream));
}

extern "C" int fb200_dense_sweep_accel(const double* A, int64_t lda, int64_t M, int64_t N, const double* xa1, int loss,
                                       const double* b, const double* za0, double c, double* za1, double* z, double* r,
                                       double* g, int bb, const double* x0, const double* xhat, const double* dx,
                                       double tau, double* scal, void* ws, size_t ws_bytes, void* stream) {
    if (!za0 || !za1) { set_error("dense_sweep_accel: za0 / za1 required"); return 1; }
    return sweep_launch(A, lda, M, N, xa1, loss, b, z, r, g, bb, x0, xhat, dx, tau, scal, ws, ws_bytes, stream, za0, za1, c);
}

// Single-pass FBS sweep, grid variant: the same contract as dense_sweep.cu --
//
//     z = A x ,  r = gradf(z) ,  f = sum loss(z) ,  g = A^T r        -- ONE read of A from HBM
//
// (reference linalg.py:41 `A @ x` at __init__.py:187 and `A.T @ r` at :248, with f / gradf between them) -- but
// without thread-block clusters.  dense_sweep.cu keeps a row on chip in the distributed shared memory of a 16-CTA
// cluster; at N = 100000 only seven such clusters fit a B200 (112 of 148 SMs; a cluster must sit inside one GPC),
// and the kernel is bound by what one SM can stream (shared-memory bandwidth: a slab is written once by the bulk
// copy and read twice), not by HBM: measured 5.4 ms = 5.9 TB/s with 7 clusters, 920 GB/s per cluster, while the
// two-pass kernels reach 7.2 TB/s on all SMs.  Here the CTAs that share a band of rows exchange their partial
// dots through L2 instead of DSMEM, so any (slabs x bands) grid works: 18 column slabs x 8 row bands = 144 SMs at
// N = 100000.
//
//   * CTA (band, rank) owns rows [band*M/bands, (band+1)*M/bands) and columns [rank*Nc, (rank+1)*Nc);
//   * producer warp: one cp.async.bulk per row slab into a ring of NST stages (full / empty mbarriers);
//   * A-group (8 warps): x slab in registers; per row: dot, block reduce, and ONE 16-byte store of the partial
//     into the band's exchange ring in global memory -- two (32-bit half, 32-bit sequence flag) pairs, so each
//     8-byte half validates itself (no fence, no second flag store);
//   * exchange warps (4, rows dealt round-robin): lane k polls rank k's entry of the row (ld.relaxed.gpu, L2), the
//     warp adds the S partials in a fixed butterfly -- every CTA of the band forms the bit-identical z_i --, lane 0
//     evaluates the loss once (r_i = gradf(z_i)) and hands r_i to the B-group through a shared-memory ring +
//     mbarrier.  One warp would bound the kernel by its own latency per row (an L2 round trip or two + the loss:
//     ~1 us, measured as a per-row time that did not shrink with the slab; 1.17 us with the logistic loss);
//   * B-group (8 warps): g slab accumulators in registers; per row re-reads the slab from shared memory:
//     g += A[i, slab] * r_i, then frees the stage;
//   * each band writes its g partial; the Barzilai-Borwein epilogue adds the band partials in index order.
//
// Flow control: a sender's A-group is at most NST rows ahead of its own B-group, which needs every CTA's partial
// of a row before it frees that row's stage, so a sender is never more than 2*NST <= 16 rows ahead of any
// receiver: the exchange ring has 32 slots per band and is never overrun.  Sequence flags are unique per row and
// per launch (a process-wide counter), so stale entries never match.  All CTAs of a band must be co-resident:
// the kernel is launched cooperatively with at most one CTA per SM.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fb200 {

int launch_bb(int bb, const double* gsrc, int nsplit, int64_t ld, int64_t n, double* g, const double* x0,
              const double* xhat, const double* dx, double tau, double* scal, Workspace& w, cudaStream_t st);

constexpr int GS_GROUP   = 256;                   // threads in the A-group and in the B-group
#ifndef FB200_GS_XWARPS
#define FB200_GS_XWARPS 1      // measured (M x N = 40000 x 100000 / 5000 x 100000): 1 warp 4.62 / 0.59 ms, 2 warps 5.59 / 0.68, 4 warps 5.62 / 0.69
#endif
constexpr int GS_XWARPS  = FB200_GS_XWARPS;       // exchange warps: warp w takes the rows it = w (mod GS_XWARPS) of the band
constexpr int GS_THREADS = 2 * GS_GROUP + 32 + 32 * GS_XWARPS;     // + producer warp + exchange warps
constexpr int GS_MAXSTG  = 8;
constexpr int GS_RSLOT   = 16;                    // r_i ring between the exchange warp and the B-group (> NST)
constexpr int GS_MAXS    = 32;                    // column slabs per band (one lane of the exchange warp each)
constexpr int GS_TAIL    = 2 * 8 * 8 + GS_RSLOT * 8 + 2 * GS_XWARPS * 8 + (2 * GS_MAXSTG + GS_RSLOT) * 8;    // redA + r ring + loss partials + barriers
constexpr int GS_SMEM_MAX = 227 * 1024;

// gradient and value of the loss at one row, separately: the B-group waits for r_i = gradf(z_i), nobody waits for f
template <int LOSS>
__device__ __forceinline__ double gs_grad(double z, double b) {
    if (LOSS == FB200_LOSS_LEAST_SQUARES) return __dsub_rn(z, b);
    if (LOSS == FB200_LOSS_LOGISTIC) return __ddiv_rn(-b, __dadd_rn(1.0, exp(__dmul_rn(b, z))));
    return z;           // LOSS_NONE: "gradient" is z itself (g = A^T A x)
}
template <int LOSS>
__device__ __forceinline__ double gs_fval(double z, double b) {
    if (LOSS == FB200_LOSS_LEAST_SQUARES) { const double r = __dsub_rn(z, b); return __dmul_rn(r, r); }
    if (LOSS == FB200_LOSS_LOGISTIC) {
        const double ind = (b == 1.0) ? 1.0 : 0.0;
        return __dsub_rn(log(__dadd_rn(1.0, exp(z))), __dmul_rn(ind, z));
    }
    return 0.0;
}

// An exchange entry is {lo32(v), lo32(seq), hi32(v), hi32(seq) + 1}: each 8-byte half carries its own piece of the
// 64-bit sequence number of (launch, row), so a reader that sees both pieces sees both halves of v whatever the
// order in which the two halves of the 16-byte store land; zero-initialised memory never validates (hi32 + 1 > 0).
__device__ __forceinline__ void xchg_store(uint4* p, double v, uint64_t seq) {
    const uint32_t lo = uint32_t(__double2loint(v)), hi = uint32_t(__double2hiint(v));
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(uint32_t(seq)), "r"(hi),
                 "r"(uint32_t(seq >> 32) + 1u) : "memory");
}

__device__ __forceinline__ uint4 xchg_issue(const uint4* p) {          // the load only: its result is examined later
    uint4 w;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ bool xchg_check(const uint4& w, uint64_t seq, double& v) {
    v = __hiloint2double(int(w.z), int(w.x));
    return w.y == uint32_t(seq) && w.w == uint32_t(seq >> 32) + 1u;
}
__device__ __forceinline__ bool xchg_load(const uint4* p, uint64_t seq, double& v) { return xchg_check(xchg_issue(p), seq, v); }

struct GsArgs {
    const double* A;
    int64_t lda;
    int M, N, Nc, S, bands, nstage;
    const double *x, *b;
    double *z, *r, *gpart;
    int64_t ldg;
    double* fpart;
    const double* za0;        // FISTA mode (see dense_sweep.cu): x is the prox point, the gradient is taken at
    double* za1;              // z_i + c (z_i - za0_i); za1 receives z_i, fpart2 the loss at the prox point
    double cacc;
    double* fpart2;
    const double* skip;
    uint4* xchg;              // [bands][GS_XRING][S]
    uint64_t seq0;            // sequence number of row 0 of this launch (row it: seq0 + it), unique process-wide
};

template <int LOSS, int CPT>
__global__ void __launch_bounds__(GS_THREADS, 1) dense_gsweep_kernel(const GsArgs a) {
    if (a.skip && __ldcg(a.skip) != 0.0) return;       // speculative trial that must not run (fb200_trial_decide): whole grid
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int STAGE_BYTES = CPT * GS_GROUP * 16;
    const int nstage = a.nstage;
    double*   redA  = reinterpret_cast<double*>(smem + size_t(nstage) * STAGE_BYTES);       // [2][8]
    double*   rring = redA + 16;                                                          // [RSLOT]
    double*   fsm   = rring + GS_RSLOT;                                                   // [2][XWARPS]
    uint64_t* full  = reinterpret_cast<uint64_t*>(fsm + 2 * GS_XWARPS);
    uint64_t* empty = full + GS_MAXSTG;
    uint64_t* rfull = empty + GS_MAXSTG;                                                   // [RSLOT]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = a.S;
    const int band = int(blockIdx.x) / S, rank = int(blockIdx.x) - band * S;
    const int M = a.M, N = a.N, Nc = a.Nc;
    const int row_lo = int((int64_t(band) * M) / a.bands), row_hi = int((int64_t(band + 1) * M) / a.bands);
    const int c0 = rank * Nc;                                    // first column of this CTA's slab
    int ncols = N - c0;
    ncols = ncols < 0 ? 0 : (ncols > Nc ? Nc : ncols);           // even (N and Nc are even)
    const uint32_t slab_bytes = uint32_t(ncols) * 8u;
    uint4* xband = a.xchg + size_t(band) * GS_XRING * S;

    // zero the stage ring once: the tail beyond `ncols` is never written by the bulk copies
    for (int i = tid; i < (nstage * STAGE_BYTES) / 16; i += GS_THREADS)
        reinterpret_cast<double2*>(smem)[i] = make_double2(0.0, 0.0);
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GS_GROUP / 32);     // one arrival per B-group warp
        }
        for (int s = 0; s < GS_RSLOT; ++s) mbar_init(&rfull[s], 1);
        mbar_fence_init();
    }
    fence_proxy_async();        // generic-proxy zero fill before async-proxy (bulk copy) writes
    __syncthreads();

    if (warp == 2 * GS_GROUP / 32) {
        // ===================================== producer =====================================
        if (lane == 0 && slab_bytes > 0) {
            const uint64_t pol = policy_evict_first();
            int stage = 0;
            uint32_t phase = 0;
            for (int row = row_lo; row < row_hi; ++row) {
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_expect_tx(&full[stage], slab_bytes);
                bulk_load(smem + size_t(stage) * STAGE_BYTES, a.A + int64_t(row) * a.lda + c0, slab_bytes, &full[stage], pol);
                if (++stage == nstage) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp > 2 * GS_GROUP / 32) {
        // ===================================== exchange warps: z_i, loss, r_i =========================
        // One row costs a warp an L2 round trip or two (the poll) plus the loss; several warps keep several rows in
        // flight.  The B-group consumes the rows in order through the r ring, whichever warp delivers them.
        const int xw = warp - (2 * GS_GROUP / 32 + 1);
        const double* b = a.b;
        const double* za0 = a.za0;
        // loss VALUES are off the critical path: lane (k mod 32) keeps the k-th row of this warp and every 32 rows all
        // lanes evaluate theirs in one pass (the logistic log(1 + e^z) costs ~0.25 us if done serially per row)
        double facc = 0.0, facc2 = 0.0, zkeep = 0.0, zikeep = 0.0, bkeep = 0.0;
        bool kept = false;
        int k = 0;
        // logistic loss only: the first poll of a row is issued one row early, so that its L2 round trip overlaps the
        // evaluation of the previous row's gradient (exp + divide) instead of following it -- measured 2.70 -> 2.47 ms at
        // 100000 x 20000; with the one-subtraction least-squares gradient the early poll only adds a wasted load per row
        // (4.54 -> 4.99 ms at 40000 x 100000), so it is compiled out there
        constexpr bool PREFETCH = (LOSS == FB200_LOSS_LOGISTIC);
        uint4 pre = make_uint4(0u, 0u, 0u, 0u);
        if (PREFETCH && row_lo + xw < row_hi && lane < S) pre = xchg_issue(xband + size_t(xw & (GS_XRING - 1)) * S + lane);
        for (int row = row_lo + xw; row < row_hi; row += GS_XWARPS, ++k) {
            const int it = row - row_lo;
            const uint64_t flag = a.seq0 + uint64_t(it);
            const uint4* src = xband + size_t(it & (GS_XRING - 1)) * S + lane;
            const double bi = b ? __ldg(b + row) : 0.0;          // issued before the poll: off the critical path
            const double qi = za0 ? __ldg(za0 + row) : 0.0;
            double v = 0.0;
            bool ok = lane >= S || (PREFETCH && xchg_check(pre, flag, v));
            while (true) {
                if (PREFETCH && __all_sync(0xffffffffu, ok)) break;
                if (!ok) ok = xchg_load(src, flag, v);
                if (!PREFETCH && __all_sync(0xffffffffu, ok)) break;
            }
            if (PREFETCH && row + GS_XWARPS < row_hi && lane < S)
                pre = xchg_issue(xband + size_t((it + GS_XWARPS) & (GS_XRING - 1)) * S + lane);
            if (lane >= S) v = 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);    // same tree on every CTA of the band
            const double zi = v;
            const double ze = za0 ? __dadd_rn(zi, __dmul_rn(a.cacc, __dsub_rn(zi, qi))) : zi;   // p + c*(p - za0), as accel_step
            if (lane == 0) {
                const double ri = gs_grad<LOSS>(ze, bi);
                const int slot = it & (GS_RSLOT - 1);
                rring[slot] = ri;
                mbar_arrive(&rfull[slot]);                     // release: the B-group's wait orders the read of rring
                if (rank == 0) {
                    a.z[row] = ze;
                    if (LOSS != FB200_LOSS_NONE) a.r[row] = ri;
                    if (za0) a.za1[row] = zi;                  // the prox image itself
                }
            }
            if (rank == 0 && LOSS != FB200_LOSS_NONE) {
                if (lane == (k & 31)) { zkeep = ze; zikeep = zi; bkeep = bi; kept = true; }
                if ((k & 31) == 31) {
                    if (kept) {
                        facc = __dadd_rn(facc, gs_fval<LOSS>(zkeep, bkeep));
                        if (za0) facc2 = __dadd_rn(facc2, gs_fval<LOSS>(zikeep, bkeep));   // f at the prox point (line search)
                        kept = false;
                    }
                }
            }
        }
        if (rank == 0 && LOSS != FB200_LOSS_NONE) {
            if (kept) {
                facc = __dadd_rn(facc, gs_fval<LOSS>(zkeep, bkeep));
                if (za0) facc2 = __dadd_rn(facc2, gs_fval<LOSS>(zikeep, bkeep));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {                   // fixed tree over the lanes
                facc = __dadd_rn(facc, __shfl_xor_sync(0xffffffffu, facc, o));
                facc2 = __dadd_rn(facc2, __shfl_xor_sync(0xffffffffu, facc2, o));
            }
        }
        if (rank == 0) {                                       // the warps' loss partials, added in warp order
            if (lane == 0) { fsm[xw] = facc; fsm[GS_XWARPS + xw] = facc2; }
            asm volatile("bar.sync 2, %0;" ::"n"(32 * GS_XWARPS) : "memory");
            if (xw == 0 && lane == 0) {
                double f = 0.0, f2 = 0.0;
                for (int k = 0; k < GS_XWARPS; ++k) { f = __dadd_rn(f, fsm[k]); f2 = __dadd_rn(f2, fsm[GS_XWARPS + k]); }
                a.fpart[band] = f;
                if (za0) a.fpart2[band] = f2;
            }
        }
    } else if (warp < GS_GROUP / 32) {
        // ===================================== A-group: partial z = slab . x ==========================
        double2 xr[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = c0 + 2 * (tid + GS_GROUP * k);
            xr[k] = (col < c0 + ncols) ? *reinterpret_cast<const double2*>(a.x + col) : make_double2(0.0, 0.0);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int row = row_lo; row < row_hi; ++row) {
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            if (slab_bytes > 0) {
                mbar_wait(&full[stage], phase);
                const double2* s = reinterpret_cast<const double2*>(smem + size_t(stage) * STAGE_BYTES) + tid;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const double2 v = s[GS_GROUP * k];
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
            const int it = row - row_lo;
            const int par = it & 1;
            if (lane == 0) redA[par * 8 + warp] = p;
            asm volatile("bar.sync 1, %0;" ::"n"(GS_GROUP) : "memory");
            if (warp == 0) {
                double v = redA[par * 8 + (lane & 7)];
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // the 8 warp partials, fixed order
                if (lane == 0) xchg_store(xband + size_t(it & (GS_XRING - 1)) * S + rank, v, a.seq0 + uint64_t(it));
            }
            if (++stage == nstage) { stage = 0; phase ^= 1u; }
        }
    } else {
        // ===================================== B-group: g += slab * r_i ===============================
        const int t = tid - GS_GROUP;
        double2 gr[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) gr[k] = make_double2(0.0, 0.0);
        int stage = 0;
        uint32_t phase = 0;
        for (int row = row_lo; row < row_hi; ++row) {
            const int it = row - row_lo;
            const int slot = it & (GS_RSLOT - 1);
            mbar_wait(&rfull[slot], uint32_t(it / GS_RSLOT) & 1u);
            const double ri = rring[slot];
            if (slab_bytes > 0) {
                mbar_wait(&full[stage], phase);      // already complete (the A-group saw it); orders the bulk-copy data for us
                const double2* s = reinterpret_cast<const double2*>(smem + size_t(stage) * STAGE_BYTES) + t;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const double2 v = s[GS_GROUP * k];
                    gr[k].x = fma(v.x, ri, gr[k].x);
                    gr[k].y = fma(v.y, ri, gr[k].y);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
            if (++stage == nstage) { stage = 0; phase ^= 1u; }
        }
        double* gp = a.gpart + int64_t(band) * a.ldg;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = c0 + 2 * (t + GS_GROUP * k);
            if (col < c0 + ncols) *reinterpret_cast<double2*>(gp + col) = gr[k];
        }
    }
}

// sum of the per-band loss partials in index order -> scal[S_F]
__global__ void gsweep_fsum_kernel(const double* __restrict__ fpart, int n, double* out, const double* skip) {
    if (skip && __ldcg(skip) != 0.0) return;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += fpart[i];
        *out = s;
    }
}

struct GsPlan {
    bool ok;
    int S, cpt, nc, nstage, bands, smem;
};

typedef void (*GsKernel)(const GsArgs);

template <int LOSS>
static GsKernel gs_pick(int cpt) {
    switch (cpt) {
        case 4: return dense_gsweep_kernel<LOSS, 4>;
        case 6: return dense_gsweep_kernel<LOSS, 6>;
        case 8: return dense_gsweep_kernel<LOSS, 8>;
        case 10: return dense_gsweep_kernel<LOSS, 10>;
        case 11: return dense_gsweep_kernel<LOSS, 11>;
        case 12: return dense_gsweep_kernel<LOSS, 12>;
        default: return dense_gsweep_kernel<LOSS, 13>;
    }
}

static GsKernel gs_kernel_for(int loss, int cpt) {
    switch (loss) {
        case FB200_LOSS_LEAST_SQUARES: return gs_pick<FB200_LOSS_LEAST_SQUARES>(cpt);
        case FB200_LOSS_LOGISTIC: return gs_pick<FB200_LOSS_LOGISTIC>(cpt);
        default: return gs_pick<FB200_LOSS_NONE>(cpt);
    }
}

static const int GS_CPTS[7] = {4, 6, 8, 10, 11, 12, 13};

// (slabs, bands): cost of a CTA per row ~ (fixed per-row work + CPT slab pieces); minimise rows-per-band x that
static GsPlan gs_make_plan(int64_t M, int64_t N, int nsm) {
    GsPlan best{};
    best.ok = false;
    double best_cost = 0.0;
    int force_s = 0;
    if (const char* e = getenv("FB200_GSWEEP_S")) force_s = atoi(e);       // experiments only
    for (int S = 1; S <= GS_MAXS && S <= nsm; ++S) {
        if (force_s > 0 && S != force_s) continue;
        const int64_t nc = round_up((N + S - 1) / S, 2);
        int cpt = 0;
        for (int c : GS_CPTS)
            if (nc <= int64_t(GS_GROUP) * 2 * c) { cpt = c; break; }
        if (!cpt) continue;
        const int stage_bytes = cpt * GS_GROUP * 16;
        int nst = (GS_SMEM_MAX - GS_TAIL - 256) / stage_bytes;
        if (nst > GS_MAXSTG) nst = GS_MAXSTG;
        if (nst < 3) continue;
        int64_t bands = nsm / S;
        // short matrices: every band costs the epilogue one more partial vector to add (measured at 200 x 1000: 148 bands
        // of one or two rows made the 1000-element band sum a 46 us kernel), so keep at least 32 rows per band
        const int64_t by_rows = M / 32 > 0 ? M / 32 : 1;
        if (bands > by_rows) bands = by_rows;
        if (bands < 1) continue;
        const double cost = double((M + bands - 1) / bands) * (10.0 + cpt);      // measured at N = 100000: S = 16 / 18 / 21 -> 4.56 / 4.77 / 5.05 ms
        if (!best.ok || cost < best_cost) {
            best.ok = true;
            best_cost = cost;
            best.S = S; best.cpt = cpt; best.nc = int(nc); best.nstage = nst; best.bands = int(bands);
            best.smem = nst * stage_bytes + GS_TAIL;
        }
    }
    return best;
}

struct GsDeviceState {
    bool attr_done[3][16];
    int nsm;
};

static std::mutex g_gs_mutex;

// per-device: cudaFuncSetAttribute applies to the current device's context only
static GsDeviceState* gs_device_state() {
    static GsDeviceState st[64];
    static bool init[64];
    std::lock_guard<std::mutex> lock(g_gs_mutex);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); dev = 0; }
    if (!init[dev]) {
        st[dev] = GsDeviceState{};
        if (cudaDeviceGetAttribute(&st[dev].nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || st[dev].nsm <= 0) {
            cudaGetLastError();
            st[dev].nsm = 148;
        }
        init[dev] = true;
    }
    return &st[dev];
}

bool gsweep_eligible(const double* A, int64_t lda, int64_t M, int64_t N) {
    return (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % 2 == 0) && (N % 2 == 0) && lda >= N && M > 0 && N > 0 &&
           M < (int64_t(1) << 31) && N <= int64_t(GS_MAXS) * GS_GROUP * 2 * 13;
}

static std::atomic<uint64_t> g_seq{1};

// za0 != nullptr: FISTA mode; S_F then holds f at the prox point and S_AUX3 f at the extrapolated z
// raw != nullptr: leave the band partials (gradient: w.dense[band][ldg], loss: w.fpart / w.fpart + FPART_MAX / 2) to the
// caller's own epilogue (the row-sharded exchange kernel) and report their layout: raw[0] = bands, raw[1] = ldg
int gsweep_launch(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss, const double* b,
                  double* z, double* r, double* g, int bb, const double* x0, const double* xhat, const double* dx,
                  double tau, double* scal, void* ws, size_t ws_bytes, void* stream, const double* za0,
                  double* za1, double c, int64_t* raw) {
    if (!gsweep_eligible(A, lda, M, N) || reinterpret_cast<uintptr_t>(x) % 16 != 0) {
        set_error("dense_gsweep: matrix not eligible (needs 16-byte aligned base and x, even lda and N)");
        return 1;
    }
    if (ws_bytes < fb200_workspace_bytes(M, N)) { set_error("dense_gsweep: workspace too small"); return 1; }
    GsDeviceState* ds = gs_device_state();
    GsPlan p = gs_make_plan(M, N, ds->nsm);
    if (!p.ok) { set_error("dense_gsweep: no feasible grid for N=%lld", (long long)N); return 1; }
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t ldg = round_up(N, 256);
    const int64_t cap = int64_t(fb200_workspace_bytes(M, N) - DENSE_OFF) / 8 / ldg;
    if (p.bands > cap) p.bands = int(cap);
    const int fcap = za0 ? FPART_MAX / 2 : FPART_MAX;       // FISTA mode keeps two loss partials per band
    if (p.bands > fcap) p.bands = fcap;
    if (p.bands * p.S > GS_XCTAS) p.bands = GS_XCTAS / p.S;
    if (p.bands < 1) { set_error("dense_gsweep: workspace admits no band"); return 1; }
    GsKernel k = gs_kernel_for(loss, p.cpt);
    const int li = loss == FB200_LOSS_LEAST_SQUARES ? 1 : (loss == FB200_LOSS_LOGISTIC ? 2 : 0);
    {
        std::lock_guard<std::mutex> lock(g_gs_mutex);
        if (!ds->attr_done[li][p.cpt]) {
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_MAX) != cudaSuccess) {
                set_error("dense_gsweep: cannot raise the shared-memory limit: %s", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
            ds->attr_done[li][p.cpt] = true;
        }
    }
    const int64_t rows_max = (M + p.bands - 1) / p.bands + 1;
    const uint64_t s0 = g_seq.fetch_add(uint64_t(rows_max));
    GsArgs a{};
    a.A = A; a.lda = lda; a.M = int(M); a.N = int(N); a.Nc = p.nc; a.S = p.S; a.bands = p.bands; a.nstage = p.nstage;
    a.x = x; a.b = b; a.z = z; a.r = r; a.gpart = w.dense; a.ldg = ldg; a.fpart = w.fpart;
    a.za0 = za0; a.za1 = za1; a.cacc = c; a.fpart2 = w.fpart + FPART_MAX / 2;
    a.skip = isnan(tau) ? scal + FB200_S_SKIP : nullptr;      // speculative trial: see fb200_trial_decide
    a.xchg = reinterpret_cast<uint4*>(w.xchg);
    a.seq0 = s0;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(unsigned(p.bands * p.S));
    cfg.blockDim = dim3(GS_THREADS);
    cfg.dynamicSmemBytes = size_t(p.smem);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeCooperative;      // every CTA of a band must be resident: they wait for each other
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, a);
    if (e != cudaSuccess) { set_error("dense_gsweep: launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    if (raw) {
        raw[0] = p.bands;
        raw[1] = ldg;
        return 0;
    }
    if (loss != FB200_LOSS_NONE) {
        if (za0) {
            gsweep_fsum_kernel<<<1, 32, 0, st>>>(a.fpart2, p.bands, scal + FB200_S_F, a.skip);       // prox point: the line-search value
            gsweep_fsum_kernel<<<1, 32, 0, st>>>(a.fpart, p.bands, scal + FB200_S_AUX3, a.skip);     // extrapolated point
        } else {
            gsweep_fsum_kernel<<<1, 32, 0, st>>>(a.fpart, p.bands, scal + FB200_S_F, a.skip);
        }
        if (check_launch("gsweep_fsum_kernel")) return 1;
    }
    // Barzilai-Borwein epilogue: fixed-order sum of the band partials (+ reductions)
    if (g) return launch_bb(bb, w.dense, p.bands, ldg, N, g, x0, xhat, dx, tau, scal, w, st);
    return 0;
}

int gsweep_plan(int64_t M, int64_t N, int* plan) {
    GsPlan p = gs_make_plan(M, N, gs_device_state()->nsm);
    if (!p.ok) return 1;
    plan[0] = p.S; plan[1] = p.bands; plan[2] = p.nstage; plan[3] = p.nc; plan[4] = p.cpt;
    return 0;
}

}  // namespace fb200

// K4 / K7: the two HBM-bound contractions of the FBS loop over a dense row-major fp64 matrix,
//     z = A x      (reference linalg.py:41 `A @ x`,   called at __init__.py:187,212)
//     g = A^T r    (reference linalg.py:41 `A.T @ x`, called at __init__.py:248)
//
// B200 design.  A is streamed exactly once per call by a persistent, warp-specialised kernel:
//   * one CTA per SM, 1 producer warp + 8 consumer warps;
//   * the producer issues TMA tile loads (cp.async.bulk.tensor.2d, 16 rows x 256 cols = 32 KB,
//     L2 evict-first) into a 6-stage shared-memory ring guarded by full/empty mbarriers, so
//     ~190 KB per SM (~28 MB chip-wide) are in flight without costing registers;
//   * the vector operand tile (x for A x, r for A^T r) rides in the same stage via a 1-D TMA
//     (out-of-bounds elements are zero-filled by the hardware, so ragged edges need no masks);
//   * consumers read the tile with conflict-free 128-bit LDS and accumulate with DFMA.
// Work decomposition is a static split: for A x an item is (16-row block, column chunk), for
// A^T r it is (256-column tile, row chunk); the chunk count S is chosen on the host so that the
// number of items is a near multiple of the CTA count (tail < 4%).  Each item writes one partial
// to the workspace; the fused epilogue kernel (loss / Barzilai-Borwein, vector_kernels.cu) adds
// the S partials in index order -- no atomics, bit-reproducible.
//
// A matrix whose base or leading dimension is not 16-byte aligned cannot be described by a TMA
// tensor map; it takes the plain-load kernels at the bottom (same partial layout).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fb200 {

int launch_loss(int loss, const double* zsrc, int nsplit, int64_t ld, const double* b, int64_t m, double* z,
                double* r, double* scal, Workspace& w, cudaStream_t st);
int launch_bb(int bb, const double* gsrc, int nsplit, int64_t ld, int64_t n, double* g, const double* x0,
              const double* xhat, const double* dx, double tau, double* scal, Workspace& w, cudaStream_t st);

constexpr int TR      = 16;    // tile rows
constexpr int TC      = 256;   // tile cols (TMA box limit per dimension)
constexpr int NSTAGE  = 6;
constexpr int NCONS   = 8;     // consumer warps; each owns TR / NCONS = 2 rows of a tile
constexpr int THREADS = (NCONS + 1) * 32;
constexpr int A_TILE_BYTES = TR * TC * 8;            // 32768
constexpr int V_TILE_BYTES = TC * 8;                 // 2048 reserved per stage (x: 256, r: 16 used)
constexpr int STAGE_BYTES  = A_TILE_BYTES + V_TILE_BYTES;
constexpr int RED_SMEM     = NCONS * TC * 8;         // 16384, A^T r cross-warp reduction
constexpr int SMEM_BYTES   = NSTAGE * STAGE_BYTES + RED_SMEM + 2 * NSTAGE * 8 + 128;

struct StreamPlan {
    int nrb;      // row blocks   = ceil(M / TR)
    int nct;      // column tiles = ceil(N / TC)
    int nsplit;   // S
    int items;    // MODE 0: nrb * S ; MODE 1: nct * S
    int grid;
    int64_t ld;   // leading dimension of the partial buffer (elements)
};

// MODE 0: partial[s][rb*16 + row] = sum over the item's column tiles of A[row, :] . x
// MODE 1: partial[s][ct*256 + col] = sum over the item's row blocks of A[:, col] . r
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
dense_stream_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapV,
                    double* __restrict__ partial, StreamPlan plan) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double*   a_tiles = reinterpret_cast<double*>(smem);
    double*   v_tiles = reinterpret_cast<double*>(smem + NSTAGE * A_TILE_BYTES);
    double*   red     = reinterpret_cast<double*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* full    = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES + RED_SMEM);
    uint64_t* empty   = full + NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCONS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int outer_n = (MODE == 0) ? plan.nrb : plan.nct;   // items are (outer, split)
    const int inner_n = (MODE == 0) ? plan.nct : plan.nrb;   // tiles walked inside an item
    (void)outer_n;
    const int q_begin = int((int64_t(blockIdx.x) * plan.items) / gridDim.x);
    const int q_end   = int((int64_t(blockIdx.x + 1) * plan.items) / gridDim.x);

    if (warp == NCONS) {
        // ===================== producer warp: one elected lane drives TMA =====================
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first();   // A is read once per call
            const uint64_t pol_keep   = policy_evict_last();    // the vector is re-read by every CTA
            int stage = 0;
            uint32_t phase = 0;
            for (int q = q_begin; q < q_end; ++q) {
                const int outer = q / plan.nsplit, s = q - outer * plan.nsplit;
                const int i_lo = int((int64_t(s) * inner_n) / plan.nsplit);
                const int i_hi = int((int64_t(s + 1) * inner_n) / plan.nsplit);
                for (int i = i_lo; i < i_hi; ++i) {
                    const int rb = (MODE == 0) ? outer : i;
                    const int ct = (MODE == 0) ? i : outer;
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], A_TILE_BYTES + (MODE == 0 ? TC * 8 : TR * 8));
                    tma_load_2d(a_tiles + size_t(stage) * TR * TC, &mapA, ct * TC, rb * TR, &full[stage], pol_stream);
                    tma_load_1d(v_tiles + size_t(stage) * TC, &mapV, (MODE == 0) ? ct * TC : rb * TR, &full[stage], pol_keep);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ============================== consumer warps ==============================================
    int stage = 0;
    uint32_t phase = 0;
    const int row0 = 2 * warp;   // this warp's two rows inside every tile
    for (int q = q_begin; q < q_end; ++q) {
        const int outer = q / plan.nsplit, s = q - outer * plan.nsplit;
        const int i_lo = int((int64_t(s) * inner_n) / plan.nsplit);
        const int i_hi = int((int64_t(s + 1) * inner_n) / plan.nsplit);

        if (MODE == 0) {
            double acc0 = 0.0, acc1 = 0.0;
            for (int i = i_lo; i < i_hi; ++i) {
                mbar_wait(&full[stage], phase);
                const double2* a0 = reinterpret_cast<const double2*>(a_tiles + size_t(stage) * TR * TC + size_t(row0) * TC);
                const double2* a1 = a0 + TC / 2;
                const double2* xv = reinterpret_cast<const double2*>(v_tiles + size_t(stage) * TC);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 x = xv[lane + 32 * k];
                    const double2 p = a0[lane + 32 * k];
                    const double2 r = a1[lane + 32 * k];
                    acc0 = fma(p.x, x.x, acc0);
                    acc0 = fma(p.y, x.y, acc0);
                    acc1 = fma(r.x, x.x, acc1);
                    acc1 = fma(r.y, x.y, acc1);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            }
            acc0 = warp_sum(acc0);
            acc1 = warp_sum(acc1);
            if (lane == 0) {
                double* dst = partial + int64_t(s) * plan.ld + int64_t(outer) * TR + row0;
                dst[0] = acc0;
                dst[1] = acc1;
            }
        } else {
            double acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.0;
            for (int i = i_lo; i < i_hi; ++i) {
                mbar_wait(&full[stage], phase);
                const double2* a0 = reinterpret_cast<const double2*>(a_tiles + size_t(stage) * TR * TC + size_t(row0) * TC);
                const double2* a1 = a0 + TC / 2;
                const double   r0 = v_tiles[size_t(stage) * TC + row0];
                const double   r1 = v_tiles[size_t(stage) * TC + row0 + 1];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 p = a0[lane + 32 * k];
                    const double2 t = a1[lane + 32 * k];
                    acc[2 * k]     = fma(p.x, r0, acc[2 * k]);
                    acc[2 * k + 1] = fma(p.y, r0, acc[2 * k + 1]);
                    acc[2 * k]     = fma(t.x, r1, acc[2 * k]);
                    acc[2 * k + 1] = fma(t.y, r1, acc[2 * k + 1]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            }
            // cross-warp reduction of the 8 row-pair partials, fixed order
            double2* rw = reinterpret_cast<double2*>(red + size_t(warp) * TC);
#pragma unroll
            for (int k = 0; k < 4; ++k) rw[lane + 32 * k] = make_double2(acc[2 * k], acc[2 * k + 1]);
            asm volatile("bar.sync 1, %0;" ::"n"(NCONS * 32) : "memory");
            {
                const int col = threadIdx.x;   // 0..255
                double sum = red[col];
#pragma unroll
                for (int w = 1; w < NCONS; ++w) sum += red[w * TC + col];
                partial[int64_t(s) * plan.ld + int64_t(outer) * TC + col] = sum;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCONS * 32) : "memory");
        }
    }
}

// ---- plain-load kernels for matrices TMA cannot describe (odd lda / unaligned base) --------------
// z = A x : one warp per row, lanes stride the columns.
__global__ void __launch_bounds__(256)
gemv_plain_kernel(const double* __restrict__ A, int64_t lda, int64_t M, int64_t N, const double* __restrict__ x,
                  double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw  = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = wid; row < M; row += nw) {
        const double* a = A + row * lda;
        double acc = 0.0;
        for (int64_t j = lane; j < N; j += 32) acc = fma(a[j], __ldg(&x[j]), acc);
        acc = warp_sum(acc);
        if (lane == 0) partial[row] = acc;
    }
}

// g = A^T r : block = 256 columns x one row chunk; partial[s][col]
__global__ void __launch_bounds__(256)
gemvT_plain_kernel(const double* __restrict__ A, int64_t lda, int64_t M, int64_t N, const double* __restrict__ r,
                   double* __restrict__ partial, int nsplit, int64_t ld) {
    const int64_t col = int64_t(blockIdx.x) * 256 + threadIdx.x;
    const int s = blockIdx.y;
    const int64_t r_lo = (int64_t(s) * M) / nsplit, r_hi = (int64_t(s + 1) * M) / nsplit;
    if (col >= N) return;
    double acc = 0.0;
    for (int64_t i = r_lo; i < r_hi; ++i) acc = fma(A[i * lda + col], __ldg(&r[i]), acc);
    partial[int64_t(s) * ld + col] = acc;
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map_2d(CUtensorMap* map, const double* A, int64_t lda, int64_t M, int64_t N) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return 1; }
    cuuint64_t dims[2]    = {cuuint64_t(N), cuuint64_t(M)};
    cuuint64_t strides[1] = {cuuint64_t(lda) * 8};
    cuuint32_t box[2]     = {TC, TR};
    cuuint32_t estr[2]    = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(A), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d, M=%lld N=%lld lda=%lld) failed: %d", (long long)M, (long long)N, (long long)lda, int(rc)); return 1; }
    return 0;
}

static int make_map_1d(CUtensorMap* map, const double* v, int64_t n, int box_elems) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return 1; }
    cuuint64_t dims[1] = {cuuint64_t(n)};
    cuuint64_t strides[1] = {0};
    cuuint32_t box[1]  = {cuuint32_t(box_elems)};
    cuuint32_t estr[1] = {1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, const_cast<double*>(v), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(1d, n=%lld) failed: %d", (long long)n, int(rc)); return 1; }
    return 0;
}

static bool tma_ok(const double* A, int64_t lda, int64_t M, int64_t N) {
    return (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % 2 == 0) && lda >= N && M > 0 && N > 0 &&
           M < (int64_t(1) << 31) && N < (int64_t(1) << 31);
}

// choose the split so that items fill the grid with < 4% tail, keeping >= 4 tiles per item
static StreamPlan make_plan(int mode, int64_t M, int64_t N) {
    StreamPlan p;
    p.nrb = int((M + TR - 1) / TR);
    p.nct = int((N + TC - 1) / TC);
    const int outer = mode == 0 ? p.nrb : p.nct;
    const int inner = mode == 0 ? p.nct : p.nrb;
    int G = sm_count();
    if (const char* e = getenv("FB200_STREAM_GRID")) { int v = atoi(e); if (v > 0 && v < G) G = v; }   // experiments only
    int best_s = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= MAX_SPLIT; ++s) {
        if (s > 1 && inner / s < 4) break;
        const int64_t items = int64_t(outer) * s;
        const int64_t waves = (items + G - 1) / G;
        const double eff = double(items) / double(waves * G);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_s = s; }
        if (eff >= 0.96) break;
    }
    p.nsplit = best_s;
    p.items  = outer * best_s;
    p.grid   = p.items < G ? p.items : G;
    p.ld     = mode == 0 ? round_up(M, TR) : round_up(N, TC);
    return p;
}

static int ensure_smem_attr() {
    static DeviceOnce once;
    return once.run([] {
        if (cudaFuncSetAttribute(dense_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(dense_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            set_error("cudaFuncSetAttribute(max dynamic smem %d) failed: %s", SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
        return 0;
    });
}

size_t dense_partial_elems(int64_t M, int64_t N) {
    // two-pass split partials: MAX_SPLIT x max(ldz, ldg); single-pass sweep: <= 160 cluster partials of
    // ldg doubles with (#clusters x ldg) bounded by ~148 x 6656 (dense_sweep.cu)
    const size_t two_pass = size_t(MAX_SPLIT) * size_t(round_up(M, TR) > round_up(N, TC) ? round_up(M, TR) : round_up(N, TC));
    const size_t sweep    = std::min<size_t>(size_t(160) * size_t(round_up(N, TC)), size_t(1200000));
    return two_pass > sweep ? two_pass : sweep;
}

// z-partials: returns nsplit / ld through the plan
static int run_gemv(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, double* partial,
                    int* nsplit, int64_t* ld, cudaStream_t st) {
    if (tma_ok(A, lda, M, N) && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
        if (ensure_smem_attr()) return 1;
        StreamPlan p = make_plan(0, M, N);
        CUtensorMap mapA, mapV;
        if (make_map_2d(&mapA, A, lda, M, N) || make_map_1d(&mapV, x, N, TC)) return 1;
        dense_stream_kernel<0><<<p.grid, THREADS, SMEM_BYTES, st>>>(mapA, mapV, partial, p);
        *nsplit = p.nsplit;
        *ld     = p.ld;
        return check_launch("dense_stream_kernel<Ax>");
    }
    const int grid = int(std::min<int64_t>((M + 7) / 8, int64_t(sm_count()) * 8));
    gemv_plain_kernel<<<grid, 256, 0, st>>>(A, lda, M, N, x, partial);
    *nsplit = 1;
    *ld     = round_up(M, TR);
    return check_launch("gemv_plain_kernel");
}

static int run_gemvT(const double* A, int64_t lda, int64_t M, int64_t N, const double* r, double* partial,
                     int* nsplit, int64_t* ld, cudaStream_t st) {
    if (tma_ok(A, lda, M, N) && reinterpret_cast<uintptr_t>(r) % 16 == 0) {
        if (ensure_smem_attr()) return 1;
        StreamPlan p = make_plan(1, M, N);
        CUtensorMap mapA, mapV;
        if (make_map_2d(&mapA, A, lda, M, N) || make_map_1d(&mapV, r, M, TR)) return 1;
        dense_stream_kernel<1><<<p.grid, THREADS, SMEM_BYTES, st>>>(mapA, mapV, partial, p);
        *nsplit = p.nsplit;
        *ld     = p.ld;
        return check_launch("dense_stream_kernel<ATr>");
    }
    const int nblk = int((N + 255) / 256);
    int s = int(std::min<int64_t>(MAX_SPLIT, std::max<int64_t>(1, (int64_t(sm_count()) * 4) / nblk)));
    if (s > M) s = int(M);
    gemvT_plain_kernel<<<dim3(nblk, s), 256, 0, st>>>(A, lda, M, N, r, partial, s, round_up(N, TC));
    *nsplit = s;
    *ld     = round_up(N, TC);
    return check_launch("gemvT_plain_kernel");
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_dense_uses_tma(const double* A, int64_t lda, int64_t M, int64_t N) {
    return tma_ok(A, lda, M, N) ? 1 : 0;
}

extern "C" int fb200_gemv_loss(const double* A, int64_t lda, int64_t M, int64_t N, const double* x, int loss,
                               const double* b, double* z, double* r, double* scal, void* ws, size_t ws_bytes,
                               void* stream) {
    if (M <= 0 || N <= 0 || lda < N) { set_error("gemv_loss: bad shape M=%lld N=%lld lda=%lld", (long long)M, (long long)N, (long long)lda); return 1; }
    if (ws_bytes < fb200_workspace_bytes(M, N)) { set_error("gemv_loss: workspace too small (%zu < %zu)", ws_bytes, fb200_workspace_bytes(M, N)); return 1; }
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int nsplit = 1;
    int64_t ld = 0;
    if (run_gemv(A, lda, M, N, x, w.dense, &nsplit, &ld, st)) return 1;
    return launch_loss(loss, w.dense, nsplit, ld, b, M, z, r, scal, w, st);
}

extern "C" int fb200_gemvT_bb(const double* A, int64_t lda, int64_t M, int64_t N, const double* r, double* g, int bb,
                              const double* x0, const double* xhat, const double* dx, double tau, double* scal,
                              void* ws, size_t ws_bytes, void* stream) {
    if (M <= 0 || N <= 0 || lda < N) { set_error("gemvT_bb: bad shape M=%lld N=%lld lda=%lld", (long long)M, (long long)N, (long long)lda); return 1; }
    if (ws_bytes < fb200_workspace_bytes(M, N)) { set_error("gemvT_bb: workspace too small (%zu < %zu)", ws_bytes, fb200_workspace_bytes(M, N)); return 1; }
    Workspace w(ws);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int nsplit = 1;
    int64_t ld = 0;
    if (run_gemvT(A, lda, M, N, r, w.dense, &nsplit, &ld, st)) return 1;
    return launch_bb(bb, w.dense, nsplit, ld, N, g, x0, xhat, dx, tau, scal, w, st);
}

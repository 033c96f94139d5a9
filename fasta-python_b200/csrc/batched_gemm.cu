// K14: the two contractions of a BATCH of FBS solves that share A (regularisation path / many
// right-hand sides; BASELINE config 5).  With B columns the GEMVs become GEMMs,
//
//     Z (M x B) = A (M x N) . X (N x B)            [forward]     reference linalg.py:41 `A @ x`, per column
//     G (N x B) = A^T (N x M) . R (M x B)          [adjoint]     reference linalg.py:41 `A.T @ x`, per column
//
// with A row-major fp64 and X, R, Z, G row-major with the batch index fastest.  Arithmetic intensity
// is B/4 flop/B (64 at B=256): compute-bound in fp64.  tcgen05.mma has no f64 kind (f16/tf32/f8f6f4/
// i8/mx* only), so the fp64 tensor path on sm_100a is the warp-level DMMA `mma.sync.m8n8k4.f64`;
// this kernel is a cp.async multi-stage, 128x128x16-tiled DMMA GEMM:
//   * 16 warps per CTA in a 4x4 grid, warp tile 32x32 = 16 DMMA tiles, accumulators in registers;
//   * operands staged in shared memory with padded pitches chosen so that every fragment load
//     (one double per lane) is bank-conflict free per half-warp;
//   * 4-stage cp.async ring (16-byte copies, zero-fill predication at ragged edges);
//   * each output element is one fixed-order dot product: column j of the result does not depend
//     on the other columns, and results are bit-reproducible.
// An Ozaki-style int8 split on tcgen05.mma.kind::i8 is the planned successor.
#include "common.cuh"

namespace fb200 {

constexpr int GB_BM = 128, GB_BN = 128, GB_BK = 16;
constexpr int GB_THREADS = 512;
constexpr int GB_STAGES  = 4;
constexpr int GB_PITCH_K = GB_BK + 4;      // A tile stored [m][k] (forward):  pitch 20 doubles
constexpr int GB_PITCH_N = GB_BN + 4;      // tiles stored [k][n] / [k][m]:     pitch 132 doubles
constexpr int GB_A_FWD   = GB_BM * GB_PITCH_K;     // doubles
constexpr int GB_A_ADJ   = GB_BK * GB_PITCH_N;
constexpr int GB_B_TILE  = GB_BK * GB_PITCH_N;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int sz = pred ? 16 : 0;       // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ADJ == 0:  C[m][n] = sum_k A[m*lda + k] * Bm[k*ldb + n]      (A row-major M_g x K)
// ADJ == 1:  C[m][n] = sum_k A[k*lda + m] * Bm[k*ldb + n]      (A row-major K x M_g, used transposed)
// requirements: lda, ldb, ldc even; base pointers 16-byte aligned (checked on the host).
template <int ADJ>
__global__ void __launch_bounds__(GB_THREADS, 1)
batched_gemm_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ Bm, int64_t ldb,
                    double* __restrict__ C, int64_t ldc, int64_t split_stride, int Mg, int Ng, int K) {
    extern __shared__ __align__(16) double gsm[];
    constexpr int A_TILE = ADJ ? GB_A_ADJ : GB_A_FWD;
    double* As = gsm;
    double* Bs = gsm + GB_STAGES * A_TILE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;               // 4 x 4 warps, 32 x 32 each
    const int m0 = blockIdx.y * GB_BM, n0 = blockIdx.x * GB_BN;
    // split-K: slice z of gridDim.z handles k tiles [kt_lo, kt_hi) and writes its own partial C
    const int nk_all = (K + GB_BK - 1) / GB_BK;
    const int kt_lo = int((int64_t(blockIdx.z) * nk_all) / gridDim.z);
    const int kt_hi = int((int64_t(blockIdx.z + 1) * nk_all) / gridDim.z);
    const int nk = kt_hi - kt_lo;
    C += int64_t(blockIdx.z) * split_stride;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = (kt_lo + kt) * GB_BK;
        double* as = As + stage * A_TILE;
        double* bs = Bs + stage * GB_B_TILE;
        if (!ADJ) {
            // A tile: 128 rows x 16 k  = 1024 16-byte chunks, 2 per thread
#pragma unroll
            for (int c = tid; c < GB_BM * GB_BK / 2; c += GB_THREADS) {
                const int row = c >> 3, kc = (c & 7) * 2;
                const bool ok = (m0 + row < Mg) && (k0 + kc < K);
                cp_async16(as + row * GB_PITCH_K + kc, A + int64_t(m0 + row) * lda + k0 + kc, ok);
            }
        } else {
            // A tile: 16 k-rows x 128 m = 1024 chunks
#pragma unroll
            for (int c = tid; c < GB_BK * GB_BM / 2; c += GB_THREADS) {
                const int kr = c >> 6, mc = (c & 63) * 2;
                const bool ok = (k0 + kr < K) && (m0 + mc < Mg);
                cp_async16(as + kr * GB_PITCH_N + mc, A + int64_t(k0 + kr) * lda + m0 + mc, ok);
            }
        }
#pragma unroll
        for (int c = tid; c < GB_BK * GB_BN / 2; c += GB_THREADS) {
            const int kr = c >> 6, nc = (c & 63) * 2;
            const bool ok = (k0 + kr < K) && (n0 + nc < Ng);
            cp_async16(bs + kr * GB_PITCH_N + nc, Bm + int64_t(k0 + kr) * ldb + n0 + nc, ok);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int s = 0; s < GB_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int fr = lane >> 2, fc = lane & 3;       // fragment row / column of this lane
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GB_STAGES - 2>();
        __syncthreads();
        {   // prefetch the tile GB_STAGES-1 ahead into the slot freed by the previous iteration
            const int nxt = kt + GB_STAGES - 1;
            if (nxt < nk) load_stage(nxt % GB_STAGES, nxt);
            cp_async_commit();
        }
        const double* as = As + (kt % GB_STAGES) * A_TILE;
        const double* bs = Bs + (kt % GB_STAGES) * GB_B_TILE;
#pragma unroll
        for (int kk = 0; kk < GB_BK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = wm * 32 + i * 8 + fr;
                af[i] = ADJ ? as[(kk + fc) * GB_PITCH_N + m] : as[m * GB_PITCH_K + kk + fc];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * GB_PITCH_N + wn * 32 + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // C fragment: lane holds C[fr][2*fc], C[fr][2*fc + 1] of each 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm * 32 + i * 8 + fr;
        if (m >= Mg) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn * 32 + j * 8 + 2 * fc;
            if (n < Ng) *reinterpret_cast<double2*>(C + int64_t(m) * ldc + n) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

static int gemm_smem(int adj) {
    return int(sizeof(double)) * GB_STAGES * ((adj ? GB_A_ADJ : GB_A_FWD) + GB_B_TILE);
}

}  // namespace fb200

using namespace fb200;

// number of K slices that fills the chip best for this shape (each slice writes its own partial C;
// the epilogue kernels add the slices in index order)
extern "C" int fb200_gemm_splits(int64_t Mg, int64_t Ng, int64_t K) {
    const int64_t tiles = ((Mg + GB_BM - 1) / GB_BM) * ((Ng + GB_BN - 1) / GB_BN);
    const int G = sm_count();
    int best = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= 8; ++s) {
        if (s > 1 && K / s < 64 * GB_BK) break;
        const int64_t items = tiles * s;
        const double eff = double(items) / double(((items + G - 1) / G) * G);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
        if (eff >= 0.95) break;
    }
    return best;
}

// C (Mg x Ng, ldc) = A (Mg x K, lda) . B (K x Ng, ldb)                 adjoint = 0
// C (Mg x Ng, ldc) = A^T, A stored (K x Mg, lda)       . B (K x Ng)    adjoint = 1
// splits > 1: C is [splits][split_stride] partial products over K slices.
extern "C" int fb200_gemm_f64(int adjoint, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                              int64_t ldc, int64_t Mg, int64_t Ng, int64_t K, int splits, int64_t split_stride,
                              void* stream) {
    if (Mg <= 0 || Ng <= 0 || K <= 0) { set_error("gemm_f64: bad shape"); return 1; }
    if ((lda | ldb | ldc) & 1 || (Ng & 1) || (reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) {
        set_error("gemm_f64: needs even leading dimensions / batch width and 16-byte aligned bases");
        return 1;
    }
    if (!adjoint && (K & 1)) { set_error("gemm_f64: forward product needs an even inner dimension"); return 1; }
    if (adjoint && (Mg & 1)) { set_error("gemm_f64: adjoint product needs an even output row count"); return 1; }
    static DeviceOnce attr_once[2];
    const int smem = gemm_smem(adjoint ? 1 : 0);
    if (attr_once[adjoint ? 1 : 0].run([&] {
            cudaError_t e = adjoint ? cudaFuncSetAttribute(batched_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                    : cudaFuncSetAttribute(batched_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) { set_error("gemm_f64: smem attribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
            return 0;
        }))
        return 1;
    if (splits < 1) splits = 1;
    dim3 grid(unsigned((Ng + GB_BN - 1) / GB_BN), unsigned((Mg + GB_BM - 1) / GB_BM), unsigned(splits));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (adjoint)
        batched_gemm_kernel<1><<<grid, GB_THREADS, smem, st>>>(A, lda, B, ldb, C, ldc, split_stride, int(Mg), int(Ng), int(K));
    else
        batched_gemm_kernel<0><<<grid, GB_THREADS, smem, st>>>(A, lda, B, ldb, C, ldc, split_stride, int(Mg), int(Ng), int(K));
    return check_launch("batched_gemm_kernel");
}

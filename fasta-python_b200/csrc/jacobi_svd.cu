// Singular-value soft threshold  U diag(shrink(s, t)) V  of a small dense matrix -- the reference's
// proximal.project_Lnuc_ball (fasta/proximal.py:44-55: la.svd, shrink of the singular values, U @ S @ V), used by
// examples/logistic_matrix_completion.py:42-45 -- without a library SVD: one-sided (Hestenes) Jacobi.
//
// The nv = min(M, N) shorter-count vectors of X (rows if M <= N, columns otherwise; length len = max(M, N)) are
// rotated pairwise until mutually orthogonal, W = P X', the rotations accumulated in P (nv x nv).  Then
// W = diag(sigma) V^T with sigma_k = |W_k|, X' = P^T W, and the prox is  P^T diag(max(sigma - t, 0) / sigma) W.
// One-sided Jacobi computes every singular value to high RELATIVE accuracy (better than bidiagonalisation), so the
// result agrees with LAPACK's to a few ulp of |X|.  One CTA per matrix: a round of the round-robin tournament
// gives every warp its own disjoint vector pairs, one block barrier per round; W and P live in shared memory when
// they fit (<= 200 KB) and in a caller-provided global scratch otherwise.  Fixed pair order and fixed-order dots:
// bit-reproducible run to run.
#include "common.cuh"

namespace fb200 {

constexpr int JS_THREADS = 1024;
constexpr int JS_MAX_SWEEPS = 60;
constexpr size_t JS_SMEM_DOUBLES = 25600;      // 200 KB

__device__ __forceinline__ double js_dot(const double* a, const double* b, int n, int lane) {
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(a[j], b[j], s);
    return warp_sum(s);
}

// info[0] = sweeps used, info[1] = 1 if converged
__global__ void __launch_bounds__(JS_THREADS)
jacobi_prox_nuclear_kernel(const double* __restrict__ X, int M, int N, int64_t ldx, double t, double* __restrict__ out,
                           int64_t ldo, double* __restrict__ svals, double* __restrict__ scratch, int* __restrict__ info) {
    extern __shared__ double js_sm[];
    __shared__ int s_rotated;
    const bool rows = M <= N;
    const int nv = rows ? M : N, len = rows ? N : M;
    double* W = scratch ? scratch : js_sm;               // [nv][len]
    double* P = W + size_t(nv) * len;                    // [nv][nv]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = JS_THREADS / 32;

    for (int64_t e = tid; e < int64_t(nv) * len; e += JS_THREADS) {
        const int k = int(e / len), j = int(e % len);
        W[e] = rows ? X[int64_t(k) * ldx + j] : X[int64_t(j) * ldx + k];
    }
    for (int e = tid; e < nv * nv; e += JS_THREADS) P[e] = (e / nv == e % nv) ? 1.0 : 0.0;
    __syncthreads();

    const int ne = nv + (nv & 1);                        // players of the tournament (one dummy if nv is odd)
    const double tol = 2.220446049250313e-16 * sqrt(double(len));
    int sweeps = 0, converged = (nv < 2);
    while (!converged && sweeps < JS_MAX_SWEEPS) {
        if (tid == 0) s_rotated = 0;
        __syncthreads();
        int rotated = 0;
        for (int r = 0; r < ne - 1; ++r) {
            for (int k = warp; k < ne / 2; k += nwarps) {
                int a = (k == 0) ? r : (r + k) % (ne - 1);
                int b = (k == 0) ? ne - 1 : (r - k + (ne - 1)) % (ne - 1);
                if (a >= nv || b >= nv) continue;
                if (a > b) { const int c = a; a = b; b = c; }
                double* wa = W + size_t(a) * len;
                double* wb = W + size_t(b) * len;
                const double alpha = js_dot(wa, wa, len, lane), beta = js_dot(wb, wb, len, lane);
                const double gamma = js_dot(wa, wb, len, lane);
                if (!(fabs(gamma) > tol * sqrt(alpha) * sqrt(beta))) continue;      // also skips zero vectors
                rotated = 1;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tn = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tn * tn), s = c * tn;
                for (int j = lane; j < len; j += 32) {
                    const double u = wa[j], v = wb[j];
                    wa[j] = c * u - s * v;
                    wb[j] = s * u + c * v;
                }
                double* pa = P + size_t(a) * nv;
                double* pb = P + size_t(b) * nv;
                for (int j = lane; j < nv; j += 32) {
                    const double u = pa[j], v = pb[j];
                    pa[j] = c * u - s * v;
                    pb[j] = s * u + c * v;
                }
            }
            __syncthreads();
        }
        if (rotated && lane == 0) s_rotated = 1;         // benign race: every writer stores 1
        __syncthreads();
        converged = !s_rotated;
        ++sweeps;
        __syncthreads();
    }

    // singular values and the per-vector factor max(sigma - t, 0) / sigma, kept in the diagonal-free spare: reuse js_dot
    // per warp and stash the factor in shared scalars (nv <= len, so the first nv entries of a small array suffice)
    __shared__ double s_fac[1024];
    for (int k0 = 0; k0 < nv; k0 += 1024) {
        const int cnt = (nv - k0 < 1024) ? nv - k0 : 1024;
        for (int k = warp; k < cnt; k += nwarps) {
            const double* wk = W + size_t(k0 + k) * len;
            const double sg = sqrt(js_dot(wk, wk, len, lane));
            if (lane == 0) {
                const double sh = sg - t;
                s_fac[k] = (sg > 0.0 && sh > 0.0) ? sh / sg : 0.0;
                if (svals) svals[k0 + k] = sg;
            }
        }
        __syncthreads();
        if (out) {
            // out' = P^T diag(fac) W restricted to the vectors k0 .. k0+cnt (accumulated over the k0 chunks)
            for (int64_t e = tid; e < int64_t(nv) * len; e += JS_THREADS) {
                const int a = int(e / len), j = int(e % len);
                double acc = 0.0;
                for (int k = 0; k < cnt; ++k) acc = fma(P[size_t(k0 + k) * nv + a] * s_fac[k], W[size_t(k0 + k) * len + j], acc);
                double* o = rows ? &out[int64_t(a) * ldo + j] : &out[int64_t(j) * ldo + a];
                *o = (k0 == 0) ? acc : *o + acc;
            }
        }
        __syncthreads();
    }
    if (tid == 0 && info) { info[0] = sweeps; info[1] = converged; }
}

}  // namespace fb200

using namespace fb200;

extern "C" size_t fb200_prox_nuclear_scratch_doubles(int64_t M, int64_t N) {
    const size_t nv = size_t(M < N ? M : N), len = size_t(M < N ? N : M);
    const size_t need = nv * len + nv * nv;
    return need <= JS_SMEM_DOUBLES ? 0 : need;
}

extern "C" int fb200_prox_nuclear(const double* X, int64_t M, int64_t N, int64_t ldx, double t, double* out, int64_t ldo,
                                  double* svals, double* scratch, int* info, void* stream) {
    if (M < 1 || N < 1 || M > (1 << 20) || N > (1 << 20)) { set_error("prox_nuclear: bad shape %lld x %lld", (long long)M, (long long)N); return 1; }
    if (!X || (!out && !svals)) { set_error("prox_nuclear: null pointer"); return 1; }
    const size_t need = fb200_prox_nuclear_scratch_doubles(M, N);
    if (need && !scratch) { set_error("prox_nuclear: %zu doubles of scratch needed for %lld x %lld", need, (long long)M, (long long)N); return 1; }
    const size_t nv = size_t(M < N ? M : N), len = size_t(M < N ? N : M);
    const size_t smem = need ? 0 : (nv * len + nv * nv) * sizeof(double);
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(jacobi_prox_nuclear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
            return check_launch("prox_nuclear (shared memory opt-in)");
    }
    jacobi_prox_nuclear_kernel<<<1, JS_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        X, int(M), int(N), ldx, t, out, ldo, svals, need ? scratch : nullptr, info);
    return check_launch("jacobi_prox_nuclear_kernel");
}

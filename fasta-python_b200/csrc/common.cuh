// Shared device/host helpers for libfasta_b200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/fasta_b200.h"

namespace fb200 {

// ------------------------------------------------------------------------------------------------
// workspace layout (bytes from the start of `ws`; the caller zero-initialises it once)
//   [0, 256)                       ticket counters for the "last block finalises" reductions
//   [256, 2048)                    per-cluster loss partials of the single-pass sweep (FPART_MAX doubles)
//   [2048, 2048 + RED_BYTES)       per-block reduction partials  [MAX_RED_BLOCKS][MAX_RED_K]
//   [XCHG_OFF, XCHG_OFF + XCHG_BYTES)   exchange rings of the grid sweep (dense_gsweep.cu): [bands][XCHG_RING][slabs] x 16 B,
//                                  touched by nothing else (its sequence flags must never be overwritten with data)
//   [DENSE_OFF, ...)               split partials of the dense maps: zp[S][ldz] then gp[S][ldg]
// ------------------------------------------------------------------------------------------------
constexpr int    MAX_RED_BLOCKS = 8192;
constexpr int    MAX_RED_K      = 12;
constexpr size_t CTR_BYTES      = 2048;
constexpr int    FPART_MAX      = 160;
constexpr size_t RED_BYTES      = size_t(MAX_RED_BLOCKS) * MAX_RED_K * sizeof(double);
constexpr int    GS_XRING       = 32;                  // rows a band's exchange ring holds (> 2 x the deepest stage ring)
constexpr int    GS_XCTAS       = 160;                 // CTAs (bands x slabs) the exchange area serves
constexpr size_t XCHG_OFF       = CTR_BYTES + RED_BYTES;
constexpr size_t XCHG_BYTES     = size_t(GS_XCTAS) * GS_XRING * 16;
constexpr size_t DENSE_OFF      = XCHG_OFF + XCHG_BYTES;
constexpr int    MAX_SPLIT      = 32;

constexpr int VEC_THREADS = 256;

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

struct Workspace {
    unsigned* counter;
    double*   fpart;
    double*   red;
    double*   dense;
    void*     xchg;
    __host__ explicit Workspace(void* ws)
        : counter(reinterpret_cast<unsigned*>(ws)),
          fpart(reinterpret_cast<double*>(static_cast<char*>(ws) + 256)),
          red(reinterpret_cast<double*>(static_cast<char*>(ws) + CTR_BYTES)),
          dense(reinterpret_cast<double*>(static_cast<char*>(ws) + DENSE_OFF)),
          xchg(static_cast<char*>(ws) + XCHG_OFF) {}
};

void set_error(const char* fmt, ...);
int  check_launch(const char* what);
int  sm_count();          // of the CURRENT device
int  current_device();    // cudaGetDevice, clamped to [0, MAX_DEVICES)

// One-time setup PER DEVICE (cudaFuncSetAttribute, occupancy queries and the plans derived from them belong to one
// device's context; a process may solve on several devices, and ctypes releases the GIL, so callers may race).
constexpr int MAX_DEVICES = 64;
struct DeviceOnce {
    std::mutex m;
    bool       done[MAX_DEVICES] = {};
    template <class F>
    int run(F&& f) {          // f() returns 0 on success; a failure is retried by the next caller
        const int dev = current_device();
        std::lock_guard<std::mutex> lock(m);
        if (done[dev]) return 0;
        const int rc = f();
        if (rc == 0) done[dev] = true;
        return rc;
    }
};

// grid for an n-element streaming kernel: enough blocks to fill the chip, capped so that the
// reduction partials fit the workspace.  148 SMs x 8 resident 256-thread blocks = 1184.
inline int vec_grid(int64_t n, int per_thread = 2) {
    int64_t want = (n + int64_t(VEC_THREADS) * per_thread - 1) / (int64_t(VEC_THREADS) * per_thread);
    int64_t cap  = int64_t(sm_count()) * 8;
    if (cap > MAX_RED_BLOCKS) cap = MAX_RED_BLOCKS;
    if (want < 1) want = 1;
    return int(want < cap ? want : cap);
}

// ------------------------------------------------------------------------------------------------
// deterministic reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum K per-thread values over the block (fixed order).  Result valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sm /* [K][32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();   // protect sm against a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sm[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = (lane < nwarp) ? sm[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// Grid-wide sum of K values per thread.  Every block writes its partial to red[block][K]; the
// block that draws the last ticket re-reads all partials in index order and writes the K totals
// through `out[k]` (pointers into the caller's scalar block; nullptr = discard).  Atomics are used
// only for the ticket, never for the sums, so the result does not depend on block scheduling.
template <int K>
__device__ __forceinline__ void grid_sum(double (&v)[K], double* red, unsigned* counter,
                                         double* const (&out)[K]) {
    __shared__ double sm[K * 32];
    __shared__ bool   is_last;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;   // kernels use 1-D or 2-D grids
    const unsigned nb  = gridDim.x * gridDim.y;
    block_sum<K>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[size_t(bid) * K + k] = v[k];
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        is_last    = (t == nb - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __ldcg(&red[size_t(b) * K + k]);
    }
    block_sum<K>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (out[k]) *out[k] = acc[k];
        *counter = 0u;   // re-arm for the next launch on this stream
    }
}

// The same, and tells the caller whether this block was the one that finished the sums (true for ALL its threads, after
// thread 0 has written the totals): work that needs every total -- the loop's decisions -- continues there.
template <int K>
__device__ __forceinline__ bool grid_sum_last(double (&v)[K], double* red, unsigned* counter, double* const (&out)[K]) {
    __shared__ double sm[K * 32];
    __shared__ bool   is_last;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    const unsigned nb  = gridDim.x * gridDim.y;
    block_sum<K>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[size_t(bid) * K + k] = v[k];
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        is_last    = (t == nb - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __ldcg(&red[size_t(b) * K + k]);
    }
    block_sum<K>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (out[k]) *out[k] = acc[k];
        *counter = 0u;   // re-arm for the next launch on this stream
    }
    __syncthreads();
    return true;
}

// ------------------------------------------------------------------------------------------------
// loss / prox element functions (compiled with -fmad=false: one rounding per numpy operation)
// ------------------------------------------------------------------------------------------------
template <int LOSS>
__device__ __forceinline__ void loss_elem(double z, double b, double& r, double& f) {
    if (LOSS == FB200_LOSS_LEAST_SQUARES) {
        r = z - b;          // gradf = z - b                       sparse_least_squares.py:42
        f = r * r;          // f = .5*norm(z-b)**2 (host finishes) sparse_least_squares.py:41
    } else if (LOSS == FB200_LOSS_LOGISTIC) {
        const double ind = (b == 1.0) ? 1.0 : 0.0;
        f = log(1.0 + exp(z)) - ind * z;      // sparse_logistic.py:47
        r = -b / (1.0 + exp(b * z));          // sparse_logistic.py:48
    } else {
        r = 0.0;
        f = 0.0;
    }
}

__device__ __forceinline__ double sign_np(double x) {   // np.sign: -1, 0, +1, nan
    return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : ((x == 0.0) ? 0.0 : x));
}

template <int PROX>
__device__ __forceinline__ double prox_elem(double h, double p0, double p1) {
    if (PROX == FB200_PROX_SHRINK || PROX == FB200_PROX_L1BALL) {
        return sign_np(h) * fmax(fabs(h) - p0, 0.0);     // proximal.py:67
    } else if (PROX == FB200_PROX_NONNEG) {
        return fmax(h, 0.0);                             // nn_least_squares.py:42
    } else if (PROX == FB200_PROX_BOX) {
        return fmin(fmax(h, p0), p1);                    // svm.py:71
    } else {
        return h;                                        // __init__.py:90
    }
}

}  // namespace fb200

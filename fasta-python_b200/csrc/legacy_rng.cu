// np.random.randn on the device, bit for bit.
//
// The reference estimates the Lipschitz constant from two Gaussian probes drawn from numpy's GLOBAL legacy
// RandomState (fasta/__init__.py:102-103: x1 = randn(*shape); x2 = randn(*shape)).  Drawing them on the host and
// uploading them is 3 ms of a 23 ms 8-GPU lasso solve and 1 s of a 1.1 s TV-4096^2 solve, so this file continues
// numpy's stream on the GPU instead: same MT19937 words, same 53-bit doubles, same polar method with its cached
// second deviate, same libm log (glibc_log.h), and it hands back the generator state numpy would have been left
// in, so that np.random.set_state() keeps the host in step (the next np.random call of the user sees no difference).
//
// numpy algorithms restated here (numpy is a dependency of the reference, not under /root/reference):
//   numpy/random/src/mt19937/mt19937.c       mt19937_gen, mt19937_next, mt19937_next_double
//   numpy/random/src/legacy/legacy-distributions.c   legacy_gauss
// Checked against np.random.randn itself: tests/test_gpu_rng.py (values and end state, several seeds / sizes /
// entry states), and on the CPU through oracle/np_legacy_rng.c, which compiles the same glibc_log.h.
//
// Pipeline (all on one stream, no host round trip):
//   1. mt_stream_kernel     one CTA regenerates MT19937 blocks in shared memory (word i+624 depends on words i,
//                           i+1, i+397: 227 threads make three words each, one barrier per block) and writes the
//                           TEMPERED word stream d[0..W)
//   2. polar_count_kernel   try t = words d[4t..4t+3] -> (x1, x2, r2); accepted iff 0 < r2 < 1; count per block
//   3. polar_scan_kernel    exclusive scan of the block counts
//   4. polar_emit_kernel    the k-th accepted try writes out[h+2k] = f*x2, out[h+2k+1] = f*x1 (legacy_gauss order)
//   5. rng_finish_kernel    words consumed -> block / position of the final state; the block is recovered from the
//                           tempered stream by inverting the tempering; cached deviate; status
// Compiled with -fmad=false: r2 = x1*x1 + x2*x2 must round like the host's separate multiply and add.
#include "common.cuh"
#include "glibc_log.h"

namespace fb200 {

constexpr int MT_N = 624, MT_M = 397, MT_LAG = MT_N - MT_M;      // 227
constexpr int RNG_STATE_WORDS = 628;      // key[624], pos, has_gauss, gauss (double, 8-byte aligned at word 626)
// state_out holds RNG_STATE_WORDS + 4 words: the state, a status word (0 ok, 1 not enough accepted tries), padding, tries used (2 words)
constexpr int POLAR_THREADS = 256, POLAR_CHUNKS = 4, POLAR_PER_BLOCK = POLAR_THREADS * POLAR_CHUNKS;

struct RngMeta {                  // lives at the start of the scratch buffer
    unsigned long long tries_used;
    double             gauss;     // f*x1 of the last accepted pair when it becomes the cached deviate
    unsigned           has_gauss;
    unsigned           fail;
    long long          c0;        // words of the entry block still unused (624 - pos)
    long long          wtot;      // words in the stream buffer
};

__device__ const double g_log_tab[256] = FB200_LOG_TAB;

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ uint32_t mt_untemper(uint32_t y) {
    y ^= (y >> 18);                                   // self-inverse for shifts >= 16
    y ^= (y << 15) & 0xefc60000u;                     // (y << 30) & mask & (mask << 15) == 0: one step inverts
    uint32_t t = y;                                   // y = t ^ ((t << 7) & m): recover 7 bits per step
    t = y ^ ((t << 7) & 0x9d2c5680u);
    t = y ^ ((t << 7) & 0x9d2c5680u);
    t = y ^ ((t << 7) & 0x9d2c5680u);
    t = y ^ ((t << 7) & 0x9d2c5680u);
    y = t;
    t = y;                                            // y = t ^ (t >> 11)
    t = y ^ (t >> 11);
    t = y ^ (t >> 11);
    return t;
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// new block n[0..624) from the old block o[0..624); threads 0..226 make three words each: word i + 624 depends on words
// i, i + 1, i + 397, so the "far" input of a thread's second word is its first, of its third its second -- its own
// results -- and ONE barrier per block (by the caller) suffices.  Word 623 also needs the NEW word 0 (thread 0's): it is
// recomputed from old words instead of waiting for it.  out != nullptr receives the tempered words.
__device__ __forceinline__ void mt_regen_block(const uint32_t* o, uint32_t* n, uint32_t* out, int tid) {
    if (tid < MT_LAG) {
        const uint32_t v0 = mt_twist(o[tid], o[tid + 1], o[tid + MT_M]);
        n[tid] = v0;
        const int i1 = tid + MT_LAG;
        const uint32_t v1 = mt_twist(o[i1], o[i1 + 1], v0);
        n[i1] = v1;
        const int i2 = tid + 2 * MT_LAG;
        uint32_t v2 = 0;
        if (i2 < MT_N) {
            const uint32_t nxt = (i2 == MT_N - 1) ? mt_twist(o[0], o[1], o[MT_M]) : o[i2 + 1];
            v2 = mt_twist(o[i2], nxt, v1);
            n[i2] = v2;
        }
        if (out) {
            out[tid] = mt_temper(v0);
            out[i1] = mt_temper(v1);
            if (i2 < MT_N) out[i2] = mt_temper(v2);
        }
    }
}

// One CTA.  d[0 .. c0) = tempered rest of the entry block, then nb regenerated blocks of 624 words.
__global__ void __launch_bounds__(256, 1)
mt_stream_kernel(const uint32_t* __restrict__ state, long long want_words, uint32_t* __restrict__ d, RngMeta* meta) {
    __shared__ uint32_t buf[2][MT_N];
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += blockDim.x) buf[0][i] = state[i];
    const int pos = int(state[MT_N]);
    const long long c0 = MT_N - pos;
    const long long nb = want_words > c0 ? (want_words - c0 + MT_N - 1) / MT_N : 0;
    if (tid == 0) {
        meta->c0 = c0;
        meta->wtot = c0 + nb * MT_N;
        meta->tries_used = 0ull;
        meta->gauss = 0.0;
        meta->has_gauss = 0u;
        meta->fail = 0u;
    }
    __syncthreads();
    for (int i = tid; i < c0; i += blockDim.x) d[i] = mt_temper(buf[0][pos + i]);
    uint32_t* out = d + c0;
    int cur = 0;
    for (long long b = 0; b < nb; ++b, cur ^= 1, out += MT_N) {
        mt_regen_block(buf[cur], buf[cur ^ 1], out, tid);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// The same stream from MANY thread blocks (large draws: the two 33.5M-element probes of a 4096^2 TV solve are 274k
// MT19937 blocks -- 35 ms each from one CTA).  MT19937's words satisfy a linear recurrence over GF(2) with a
// characteristic polynomial phi of degree 19937, so  y[n + J] = XOR_{j : g_j = 1} y[n + j]  with g = t^J mod phi
// (Haramoto et al. 2008): a CTA that holds 19937 + 623 consecutive words (33 blocks) produces the block J words
// further on by a 19937-tap XOR convolution.  tools/make_mt_jump.py computes g for the strides 64 * 16^level * digit
// blocks (fasta/mt19937_jump.npz, verified against numpy); segment s starts s * seg_units * 64 blocks into the
// stream and is reached with one jump per non-zero hexadecimal digit of s * seg_units.
//   y[n] = untempered word n after the entry block;  d[c0 + n] = temper(y[n]).
// ---------------------------------------------------------------------------------------------------
constexpr int MTJ_PREFIX = 33;                       // blocks a CTA holds: 33 * 624 = 20592 >= 19937 + 623 words
constexpr int MTJ_UNIT = 64;                         // blocks per stride unit of the polynomial table
constexpr int MTJ_LEVELS = 4, MTJ_DIGITS = 16;
constexpr int MTJ_SMEM = (MTJ_PREFIX * MT_N + 2 * MT_N) * 4;      // prefix + polynomial + jumped block
constexpr long long MTJ_PARALLEL_MIN_BLOCKS = 2048;  // below ~0.5M draws one CTA is as fast as the jumps

__global__ void __launch_bounds__(256, 2)
mt_parallel_stream_kernel(const uint32_t* __restrict__ state, const uint32_t* __restrict__ polys, long long want_words,
                          int seg_units, uint32_t* __restrict__ d, RngMeta* meta) {
    extern __shared__ __align__(16) uint32_t mtj_smem[];
    uint32_t* pre  = mtj_smem;                       // [33][624], flat: pre[k] = y[J + k]
    uint32_t* g    = pre + MTJ_PREFIX * MT_N;        // [624] polynomial words
    uint32_t* knew = g + MT_N;                       // [624] jumped block
    const int tid = threadIdx.x;
    const int pos = int(state[MT_N]);
    const long long c0 = MT_N - pos;
    const long long nb = want_words > c0 ? (want_words - c0 + MT_N - 1) / MT_N : 0;
    const long long seg_blocks = (long long)seg_units * MTJ_UNIT;
    const long long b_lo = (long long)blockIdx.x * seg_blocks;
    if (blockIdx.x == 0) {
        if (tid == 0) {
            meta->c0 = c0;
            meta->wtot = c0 + nb * MT_N;
            meta->tries_used = 0ull;
            meta->gauss = 0.0;
            meta->has_gauss = 0u;
            meta->fail = 0u;
        }
        for (int i = tid; i < c0; i += blockDim.x) d[i] = mt_temper(state[pos + i]);
    }
    if (b_lo >= nb) return;
    const long long b_hi = (b_lo + seg_blocks < nb) ? b_lo + seg_blocks : nb;
    // prefix of the stream: y[0 .. 33 * 624)
    for (int i = tid; i < MT_N; i += blockDim.x) knew[i] = state[i];
    __syncthreads();
    mt_regen_block(knew, pre, nullptr, tid);
    __syncthreads();
    for (int b = 1; b < MTJ_PREFIX; ++b) {
        mt_regen_block(pre + (b - 1) * MT_N, pre + b * MT_N, nullptr, tid);
        __syncthreads();
    }
    // one jump per non-zero hexadecimal digit of the segment's offset (in units of 64 blocks)
    const unsigned U = unsigned(blockIdx.x) * unsigned(seg_units);
    for (int level = 0; level < MTJ_LEVELS; ++level) {
        const unsigned digit = (U >> (4 * level)) & 15u;
        if (digit == 0) continue;                                              // uniform over the block
        const uint32_t* gp = polys + (size_t(level) * MTJ_DIGITS + digit) * MT_N;
        for (int i = tid; i < MT_N; i += blockDim.x) g[i] = gp[i];
        __syncthreads();
        const int i0 = tid, i1 = tid + 256, i2 = tid + 512;
        uint32_t a0 = 0, a1 = 0, a2 = 0;
        const bool has2 = i2 < MT_N;
        for (int jw = 0; jw < MT_N; ++jw) {
            uint32_t gw = g[jw];                                               // the same word in every thread: no divergence
            const uint32_t* base = pre + 32 * jw;
            while (gw) {
                const int j = __ffs(gw) - 1;
                gw &= gw - 1;
                a0 ^= base[i0 + j];
                a1 ^= base[i1 + j];
                if (has2) a2 ^= base[i2 + j];
            }
        }
        knew[i0] = a0;
        knew[i1] = a1;
        if (has2) knew[i2] = a2;
        __syncthreads();
        for (int i = tid; i < MT_N; i += blockDim.x) pre[i] = knew[i];
        __syncthreads();
        for (int b = 1; b < MTJ_PREFIX; ++b) {
            mt_regen_block(pre + (b - 1) * MT_N, pre + b * MT_N, nullptr, tid);
            __syncthreads();
        }
    }
    // the segment: its first 33 blocks are at hand, the rest is regenerated block by block in the prefix area
    uint32_t* out = d + c0 + b_lo * MT_N;
    const long long have = (b_hi - b_lo < MTJ_PREFIX) ? b_hi - b_lo : MTJ_PREFIX;
    for (long long k = tid; k < have * MT_N; k += blockDim.x) out[k] = mt_temper(pre[k]);
    out += have * MT_N;
    if (b_lo + have >= b_hi) return;
    for (int i = tid; i < MT_N; i += blockDim.x) knew[i] = pre[(MTJ_PREFIX - 1) * MT_N + i];
    __syncthreads();
    uint32_t* bufs[2] = {knew, pre};
    int cur = 0;
    for (long long b = b_lo + have; b < b_hi; ++b, cur ^= 1, out += MT_N) {
        mt_regen_block(bufs[cur], bufs[cur ^ 1], out, tid);
        __syncthreads();
    }
}

// try t of the stream -> the polar method's candidate point
__device__ __forceinline__ bool polar_try(const uint32_t* __restrict__ d, long long t, double& x1, double& x2, double& r2) {
    const uint4 w = *reinterpret_cast<const uint4*>(d + 4 * t);
    const double u1 = (double(int(w.x >> 5)) * 67108864.0 + double(int(w.y >> 6))) / 9007199254740992.0;
    const double u2 = (double(int(w.z >> 5)) * 67108864.0 + double(int(w.w >> 6))) / 9007199254740992.0;
    x1 = 2.0 * u1 - 1.0;
    x2 = 2.0 * u2 - 1.0;
    r2 = x1 * x1 + x2 * x2;
    return !(r2 >= 1.0 || r2 == 0.0);
}

__global__ void __launch_bounds__(POLAR_THREADS)
polar_count_kernel(const uint32_t* __restrict__ d, long long tries, unsigned* __restrict__ counts) {
    const long long base = (long long)blockIdx.x * POLAR_PER_BLOCK;
    int total = 0;
#pragma unroll
    for (int c = 0; c < POLAR_CHUNKS; ++c) {
        const long long t = base + c * POLAR_THREADS + threadIdx.x;
        double x1, x2, r2;
        const bool ok = t < tries && polar_try(d, t, x1, x2, r2);
        total += __syncthreads_count(ok);
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = unsigned(total);
}

// exclusive scan of counts[0..nblk) in place into offsets (64-bit), one CTA of 1024 threads
__global__ void __launch_bounds__(1024, 1)
polar_scan_kernel(const unsigned* __restrict__ counts, long long nblk, unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    const long long per = (nblk + 1023) / 1024;
    const long long lo = tid * per, hi = (lo + per < nblk) ? lo + per : nblk;
    unsigned long long s = 0;
    for (long long i = lo; i < hi; ++i) s += counts[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                      // Hillis-Steele inclusive scan
        const unsigned long long v = (tid >= o) ? part[tid - o] : 0ull;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned long long run = part[tid] - s;
    for (long long i = lo; i < hi; ++i) {
        offsets[i] = run;
        run += counts[i];
    }
    if (tid == 1023) offsets[nblk] = part[1023];              // grand total
}

__global__ void __launch_bounds__(POLAR_THREADS)
polar_emit_kernel(const uint32_t* __restrict__ d, long long tries, const unsigned long long* __restrict__ offsets,
                  const uint32_t* __restrict__ state, long long n, double* __restrict__ out, RngMeta* meta) {
    __shared__ double tab[256];
    __shared__ unsigned wcount[POLAR_CHUNKS][POLAR_THREADS / 32];
    tab[threadIdx.x] = g_log_tab[threadIdx.x];
    const int h = int(state[MT_N + 1]) != 0 ? 1 : 0;           // entry state holds a cached deviate: it is out[0]
    const long long fresh = n - h;                             // values that come from new pairs
    const long long pairs = fresh > 0 ? (fresh + 1) / 2 : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && h && n > 0) out[0] = *reinterpret_cast<const double*>(state + MT_N + 2);
    const long long base = (long long)blockIdx.x * POLAR_PER_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x1[POLAR_CHUNKS], x2[POLAR_CHUNKS], r2[POLAR_CHUNKS];
    unsigned ball[POLAR_CHUNKS];
    bool ok[POLAR_CHUNKS];
#pragma unroll
    for (int c = 0; c < POLAR_CHUNKS; ++c) {
        const long long t = base + c * POLAR_THREADS + threadIdx.x;
        ok[c] = t < tries && polar_try(d, t, x1[c], x2[c], r2[c]);
        ball[c] = __ballot_sync(0xffffffffu, ok[c]);
        if (lane == 0) wcount[c][warp] = __popc(ball[c]);
    }
    __syncthreads();
    unsigned long long rank0 = offsets[blockIdx.x];
#pragma unroll
    for (int c = 0; c < POLAR_CHUNKS; ++c) {
        unsigned before = 0;
        for (int w = 0; w < POLAR_THREADS / 32; ++w) {
            const unsigned v = wcount[c][w];
            before += (w < warp) ? v : 0u;
        }
        unsigned chunk_total = 0;
        for (int w = 0; w < POLAR_THREADS / 32; ++w) chunk_total += wcount[c][w];
        if (ok[c]) {
            const unsigned long long k = rank0 + before + __popc(ball[c] & ((1u << lane) - 1u));
            if ((long long)k < pairs) {
                const double f = sqrt(-2.0 * fb200_glibc_log(r2[c], tab) / r2[c]);
                const long long i0 = h + 2 * (long long)k;
                out[i0] = f * x2[c];
                if (i0 + 1 < n) out[i0 + 1] = f * x1[c];
                if ((long long)k == pairs - 1) {
                    meta->tries_used = (unsigned long long)(base + c * POLAR_THREADS + threadIdx.x) + 1ull;
                    if (i0 + 1 >= n) {                         // odd count: the second deviate stays cached
                        meta->gauss = f * x1[c];
                        meta->has_gauss = 1u;
                    }
                }
            }
        }
        rank0 += chunk_total;
    }
}

// state_out[0..628) = the legacy state after the draws, [628] = status, [630..632) = tries used
__global__ void __launch_bounds__(640, 1)
rng_finish_kernel(const uint32_t* __restrict__ state, const uint32_t* __restrict__ d, long long n,
                  const unsigned long long* __restrict__ total_accepted, RngMeta* meta, uint32_t* __restrict__ state_out) {
    const int tid = threadIdx.x;
    const int h = int(state[MT_N + 1]) != 0 ? 1 : 0;
    const long long fresh = n - h;
    const long long pairs = fresh > 0 ? (fresh + 1) / 2 : 0;
    const bool fail = pairs > 0 && (long long)(*total_accepted) < pairs;
    const long long used_words = 4 * (long long)meta->tries_used;
    const long long c0 = meta->c0;
    const int pos = int(state[MT_N]);
    if (tid < MT_N) {
        uint32_t v;
        if (fail || used_words <= c0) {
            v = state[tid];
        } else {
            const long long j = used_words - c0;               // words consumed from regenerated blocks, >= 1
            const long long q = (j - 1) / MT_N;
            v = mt_untemper(d[c0 + q * MT_N + tid]);
        }
        state_out[tid] = v;
    }
    if (tid == 0) {
        int new_pos = pos;
        if (!fail) {
            if (used_words <= c0) new_pos = pos + int(used_words);
            else { const long long j = used_words - c0; new_pos = int(j - ((j - 1) / MT_N) * MT_N); }
        }
        unsigned hg;
        double g;
        if (fail || n <= 0) { hg = unsigned(h); g = *reinterpret_cast<const double*>(state + MT_N + 2); }
        else if (fresh <= 0) { hg = 0u; g = 0.0; }                 // n == 1 served by the cached deviate
        else { hg = meta->has_gauss; g = hg ? meta->gauss : 0.0; }
        state_out[MT_N] = uint32_t(new_pos);
        state_out[MT_N + 1] = hg;
        *reinterpret_cast<double*>(state_out + MT_N + 2) = g;
        state_out[RNG_STATE_WORDS] = fail ? 1u : 0u;
        state_out[RNG_STATE_WORDS + 1] = 0u;
        *reinterpret_cast<unsigned long long*>(state_out + RNG_STATE_WORDS + 2) = meta->tries_used;
    }
}

static long long rng_tries(long long n) {
    const long long pairs = (n + 1) / 2;
    // acceptance probability pi/4; negative-binomial spread sqrt(pairs (1-p)) / p; 7 sigma + slack
    const double mean = double(pairs) * 1.2732395447351628, sd = sqrt(double(pairs) * 0.2146) / 0.7853981633974483;
    return (long long)(mean + 7.0 * sd) + 64;
}

struct RngLayout {
    long long tries, words_cap, nblk;
    size_t off_words, off_counts, off_offsets, total;
};

static RngLayout rng_layout(long long n) {
    RngLayout L;
    L.tries = rng_tries(n);
    L.words_cap = 4 * L.tries + 2 * MT_N + 8;
    L.nblk = (L.tries + POLAR_PER_BLOCK - 1) / POLAR_PER_BLOCK;
    L.off_words = 256;
    L.off_counts = L.off_words + size_t(round_up(L.words_cap * 4, 256));
    L.off_offsets = L.off_counts + size_t(round_up(L.nblk * 4, 256));
    L.total = L.off_offsets + size_t(round_up((L.nblk + 1) * 8, 256));
    return L;
}

}  // namespace fb200

using namespace fb200;

extern "C" size_t fb200_randn_scratch_bytes(int64_t n) { return n > 0 ? rng_layout(n).total : 1024; }

// n standard normals continuing the numpy legacy stream held in `state` (device: key[624], pos, has_gauss, gauss).
// state_out (device, 632 words): the state after the draws + status word [628] (1 = the planned number of candidate
// points did not yield enough accepted pairs; nothing may be used then -- probability < 1e-11) + tries used [630..632).
// jump_polys (device, [4][16][624] words from fasta/mt19937_jump.npz, or NULL): lets large draws generate the word
// stream from many thread blocks.
extern "C" int fb200_randn_legacy(const void* state, int64_t n, double* out, void* scratch, size_t scratch_bytes,
                                  void* state_out, const void* jump_polys, void* stream) {
    if (n <= 0) { set_error("randn_legacy: n must be positive"); return 1; }
    const RngLayout L = rng_layout(n);
    if (scratch_bytes < L.total) { set_error("randn_legacy: scratch too small (%zu < %zu)", scratch_bytes, L.total); return 1; }
    if (reinterpret_cast<uintptr_t>(scratch) % 256 != 0 || reinterpret_cast<uintptr_t>(state) % 8 != 0 ||
        reinterpret_cast<uintptr_t>(state_out) % 8 != 0) { set_error("randn_legacy: misaligned buffers"); return 1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(scratch);
    RngMeta* meta = reinterpret_cast<RngMeta*>(base);
    uint32_t* words = reinterpret_cast<uint32_t*>(base + L.off_words);
    unsigned* counts = reinterpret_cast<unsigned*>(base + L.off_counts);
    unsigned long long* offsets = reinterpret_cast<unsigned long long*>(base + L.off_offsets);
    const uint32_t* s = static_cast<const uint32_t*>(state);
    const long long want = 4 * L.tries;
    const long long nb_max = (want + MT_N - 1) / MT_N;                     // blocks to regenerate if the entry block is spent
    const long long units = (nb_max + MTJ_UNIT - 1) / MTJ_UNIT;
    if (jump_polys && nb_max >= MTJ_PARALLEL_MIN_BLOCKS && units < (1ll << (4 * MTJ_LEVELS))) {
        static DeviceOnce once;
        if (once.run([] {
                if (cudaFuncSetAttribute(mt_parallel_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MTJ_SMEM) != cudaSuccess) {
                    set_error("randn_legacy: shared-memory attribute: %s", cudaGetErrorString(cudaGetLastError()));
                    return 1;
                }
                return 0;
            }))
            return 1;
        const long long target = 2ll * sm_count();                         // two 85 KB CTAs are resident per SM
        long long seg_units = (units + target - 1) / target;
        if (seg_units < 1) seg_units = 1;
        const long long nseg = (units + seg_units - 1) / seg_units;
        mt_parallel_stream_kernel<<<unsigned(nseg), 256, MTJ_SMEM, st>>>(s, static_cast<const uint32_t*>(jump_polys), want,
                                                                         int(seg_units), words, meta);
    } else {
        mt_stream_kernel<<<1, 256, 0, st>>>(s, want, words, meta);
    }
    polar_count_kernel<<<unsigned(L.nblk), POLAR_THREADS, 0, st>>>(words, L.tries, counts);
    polar_scan_kernel<<<1, 1024, 0, st>>>(counts, L.nblk, offsets);
    polar_emit_kernel<<<unsigned(L.nblk), POLAR_THREADS, 0, st>>>(words, L.tries, offsets, s, n, out, meta);
    rng_finish_kernel<<<1, 640, 0, st>>>(s, words, n, offsets + L.nblk, meta, static_cast<uint32_t*>(state_out));
    return check_launch("randn_legacy");
}

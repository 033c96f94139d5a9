// K14 on the 5th-generation tensor cores: the two contractions of a BATCH of FBS solves,
//
//     Z (M x B) = A (M x N) . X (N x B)         per column the reference's `A @ x`    (linalg.py:41)
//     G (N x B) = A^T (N x M) . R (M x B)       per column the reference's `A.T @ r`  (linalg.py:41)
//
// in fp64 accuracy on tcgen05.mma.kind::i8 with int32 accumulators in tensor memory.
//
// tcgen05 has no f64 kind, so both operands are split ERROR-FREE into S = 8 signed 7-bit digit
// planes (Ozaki-style slicing): with a power-of-two scale per row of the left operand and per column
// of the right operand,
//
//     a_ik = sa_i * sum_s p_s[i][k] * 2^(-7s),      x_kj = sx_j * sum_t q_t[k][j] * 2^(-7t),
//
// where all p, q are integers in [-64, 64].  Every product p_s . q_t is an exact int8 GEMM; the pairs
// with s + t = d share the weight 2^(-7d) and accumulate into ONE int32 accumulator (8 accumulators of
// 128 x 64 int32 = all 512 TMEM columns).  Pairs with s + t > 7 are below 2^-56 of the row/column scale
// and are dropped (36 of 64 pairs remain), which leaves the result at least as accurate as an fp64 dot
// product of the same length.  The epilogue evaluates  sa_i * sx_j * sum_d acc_d 2^(-7d)  by Horner in
// fp64 (every int32 and every power of two is exact; one rounding per Horner step).
//
// Kernel structure (one 128 x 64 output tile and one K split per CTA, 192 threads):
//   warp 0 (one lane)  TMA producer: per 128-byte K block, the 8 right-operand digit tiles (64 x 128 B,
//                      double buffered) and the 8 left-operand digit tiles (128 x 128 B, through a ring),
//                      SWIZZLE_128B, completion on mbarriers;
//   warp 1 (one lane)  MMA issuer: for left digit s and each of the 4 K steps (32 bytes), tcgen05.mma against
//                      right digits t = 0..7-s with the A tile held in the A collector across them;
//                      tcgen05.commit releases the ring slot / the buffer;
//   warps 2..5         epilogue: tcgen05.ld of the 8 accumulators, Horner, scaling, fp64 stores.
// Integer accumulation is exact as long as (d+1) * K_split * 64 * 64 < 2^31, i.e. K_split <= 65535; the
// host picks the number of K splits accordingly (and to balance the 148 SMs).  Each split writes its own
// partial; the batched epilogue kernels add the partials in index order, so results are reproducible.
#include "common.cuh"
#include "ptx.cuh"

namespace fb200 {

constexpr int OZ_S       = 8;                      // digit planes
constexpr int OZ_BM      = 128;                    // output rows per CTA  (UMMA M)
constexpr int OZ_BN      = 64;                     // output columns per CTA (UMMA N)
constexpr int OZ_BK      = 128;                    // contraction bytes per pipeline block (one swizzle row)
constexpr int OZ_UK      = 32;                     // contraction bytes per tcgen05.mma (kind::i8)
constexpr int OZ_A_TILE  = OZ_BM * OZ_BK;          // 16 KB
constexpr int OZ_X_TILE  = OZ_BN * OZ_BK;          //  8 KB
constexpr int OZ_NA      = 5;                      // left-operand ring slots
constexpr int OZ_THREADS = 192;
constexpr int OZ_TMEM_COLS = 512;
constexpr int OZ_SMEM_DATA = OZ_NA * OZ_A_TILE + 2 * OZ_S * OZ_X_TILE;
constexpr int OZ_SMEM = OZ_SMEM_DATA + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int64_t OZ_MAX_KSPLIT = 65535 / OZ_BK * OZ_BK;     // exact int32 accumulation bound

// ---- tcgen05 wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32, M = 128, N = 64, K = 32.
// COLL selects the A-operand collector use: consecutive UMMAs that share the A tile read it from shared
// memory once (fill), reuse it from the collector buffer (use) and release it (lastuse); 0 = no reuse.
enum { OZ_COLL_NONE = 0, OZ_COLL_FILL = 1, OZ_COLL_USE = 2, OZ_COLL_LAST = 3 };
template <int COLL>
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
#define OZ_UMMA(QUAL)                                                                                          \
    asm volatile(                                                                                              \
        "{\n\t"                                                                                                \
        ".reg .pred p;\n\t"                                                                                    \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                            \
        "tcgen05.mma.cta_group::1.kind::i8" QUAL " [%0], %1, %2, %3, p;\n\t"                                   \
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)                                 \
        : "memory")
    if (COLL == OZ_COLL_FILL) OZ_UMMA(".collector::a::fill");
    else if (COLL == OZ_COLL_USE) OZ_UMMA(".collector::a::use");
    else if (COLL == OZ_COLL_LAST) OZ_UMMA(".collector::a::lastuse");
    else OZ_UMMA("");
#undef OZ_UMMA
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns: thread l of the warp receives columns c..c+7 of lane (base lane + l)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one lane of a converged warp (the compiler can then keep tcgen05 / TMA operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xFFFFFFFF;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// long wait of the epilogue warps: back off so that the spin does not compete with the MMA issuer for issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(512);
    }
}

__device__ __forceinline__ void tma_load_2d_plain(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// K-major operand tile in shared memory, rows of 128 bytes, SWIZZLE_128B (what the TMA box writes):
// 8-row groups are 1024 bytes apart (stride byte offset), descriptor version 1 (sm_100).
constexpr uint32_t OZ_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(const void* tile) { return ((smem_u32(tile) & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo) { return (uint64_t(OZ_DESC_HI) << 32) | lo; }
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 64, M = 128
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(OZ_BN >> 3) << 17) | (uint32_t(OZ_BM >> 4) << 24);

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapX,
                  const double* __restrict__ ascale, const double* __restrict__ xscale, double* __restrict__ C, int64_t ldc,
                  int64_t split_stride, const int* __restrict__ colmap, int Mg, int Ng, int mpad, int npad, int nkb_total,
                  int nblk_n, int nblk_m, int splits) {
    extern __shared__ uint8_t oz_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_raw) + 1023) & ~uintptr_t(1023));
    uint8_t*  smA = sm;
    uint8_t*  smX = sm + OZ_NA * OZ_A_TILE;
    uint64_t* afull  = reinterpret_cast<uint64_t*>(sm + OZ_SMEM_DATA);
    uint64_t* aempty = afull + OZ_NA;
    uint64_t* xfull  = aempty + OZ_NA;
    uint64_t* xempty = xfull + 2;
    uint64_t* tfull  = xempty + 2;
    uint32_t* tslot  = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bid = blockIdx.x;
    const int nb = bid % nblk_n; bid /= nblk_n;
    const int mb = bid % nblk_m;
    const int sp = bid / nblk_m;
    const int kb_lo = int((int64_t(sp) * nkb_total) / splits);
    const int kb_hi = int((int64_t(sp + 1) * nkb_total) / splits);
    const int nkb = kb_hi - kb_lo;
    const int m0 = mb * OZ_BM, n0 = nb * OZ_BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < OZ_NA; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
        mbar_init(tfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tslot, OZ_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;

    if (warp == 0) {
        // ===== TMA producer (whole warp waits, one elected lane issues) =====
        auto issue_x = [&](int k) {
            const int xb = k & 1;
            mbar_wait(&xempty[xb], ((k >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&xfull[xb], OZ_S * OZ_X_TILE);
#pragma unroll
                for (int t = 0; t < OZ_S; ++t)
                    tma_load_2d_plain(smX + (xb * OZ_S + t) * OZ_X_TILE, &mapX, (kb_lo + k) * OZ_BK, t * npad + n0, &xfull[xb]);
            }
            __syncwarp();
        };
        int slot = 0;
        uint32_t phase = 0;
        if (nkb > 0) issue_x(0);
        for (int kb = 0; kb < nkb; ++kb) {
            for (int s = 0; s < OZ_S; ++s) {
                if (s == 3 && kb + 1 < nkb) issue_x(kb + 1);
                mbar_wait(&aempty[slot], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&afull[slot], OZ_A_TILE);
                    tma_load_2d_plain(smA + slot * OZ_A_TILE, &mapA, (kb_lo + kb) * OZ_BK, s * mpad + m0, &afull[slot]);
                }
                __syncwarp();
                if (++slot == OZ_NA) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one elected lane issues) =====
        int slot = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int xb = kb & 1;
            mbar_wait(&xfull[xb], (kb >> 1) & 1);
            const uint32_t xlo = umma_desc_lo(smX + xb * OZ_S * OZ_X_TILE);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) {
                mbar_wait(&afull[slot], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t alo = umma_desc_lo(smA + slot * OZ_A_TILE);
                    // K step outermost: the OZ_S - s products of one K step share the A digit tile, which
                    // then comes from shared memory once (A collector) instead of once per product
#pragma unroll
                    for (int k = 0; k < OZ_BK / OZ_UK; ++k) {
                        const uint64_t adesc = umma_desc(alo + k * (OZ_UK >> 4));
                        const uint32_t acc = (kb > 0 || s > 0 || k > 0) ? 1u : 0u;
#pragma unroll
                        for (int t = 0; t < OZ_S - s; ++t) {
                            const uint64_t bdesc = umma_desc(xlo + t * (OZ_X_TILE >> 4) + k * (OZ_UK >> 4));
                            const uint32_t dcol = tmem + uint32_t(s + t) * OZ_BN;
                            if (OZ_S - s == 1) umma_i8<OZ_COLL_NONE>(dcol, adesc, bdesc, OZ_IDESC, acc);
                            else if (t == 0) umma_i8<OZ_COLL_FILL>(dcol, adesc, bdesc, OZ_IDESC, acc);
                            else if (t == OZ_S - s - 1) umma_i8<OZ_COLL_LAST>(dcol, adesc, bdesc, OZ_IDESC, acc);
                            else umma_i8<OZ_COLL_USE>(dcol, adesc, bdesc, OZ_IDESC, acc);
                        }
                    }
                    umma_commit(&aempty[slot]);
                    if (s == OZ_S - 1) umma_commit(&xempty[xb]);
                }
                __syncwarp();
                if (++slot == OZ_NA) { slot = 0; phase ^= 1; }
            }
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else if (warp >= 2) {
        // ===== epilogue: a warp may touch the TMEM lane quarter warp % 4 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const double sa = (m < Mg) ? ascale[m] : 0.0;
        double* crow = C + int64_t(sp) * split_stride + int64_t(m) * ldc;
        if (nkb > 0) {
            mbar_wait_backoff(tfull, 0);
            tc_fence_after();
        }
        const uint32_t tbase = tmem + (uint32_t(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < OZ_BN; c += 8) {
            double sum[8];
            if (nkb > 0) {
                int v[OZ_S][8];
#pragma unroll
                for (int d = 0; d < OZ_S; ++d) tmem_ld8(tbase + uint32_t(d * OZ_BN + c), v[d]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    double h = double(v[OZ_S - 1][j]);
#pragma unroll
                    for (int d = OZ_S - 2; d >= 0; --d) h = fma(h, 0.0078125, double(v[d][j]));
                    sum[j] = h;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) sum[j] = 0.0;
            }
            if (m < Mg) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = n0 + c + j;
                    if (n < Ng) {
                        const double val = (sum[j] * sa) * xscale[n];
                        if (colmap) crow[colmap[n]] = val;
                        else crow[n] = val;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, OZ_TMEM_COLS);
    }
}

// ---- digit-plane slicing ----------------------------------------------------------------------------
// scale = 2^(e-6) with 2^(e-1) <= maxabs < 2^e, so |v / scale| < 64; 0 for an all-zero row / column
__device__ __forceinline__ void oz_scale_of(double maxabs, double& scale, double& inv) {
    if (!(maxabs > 0.0) || !isfinite(maxabs)) { scale = 0.0; inv = 0.0; return; }
    const int e = ilogb(maxabs);
    scale = scalbn(1.0, e - 5);
    inv   = scalbn(1.0, 5 - e);
}
// the 8 signed digits of v * inv (|.| < 64): digit 0 = nearest integer, digit s = nearest integer of the
// remainder * 128^s; every step is exact in fp64
__device__ __forceinline__ void oz_digits(double v, double inv, signed char (&q)[OZ_S]) {
    double r = v * inv;
#pragma unroll
    for (int s = 0; s < OZ_S; ++s) {
        const double d = rint(r);
        q[s] = static_cast<signed char>(static_cast<int>(d));
        r = (r - d) * 128.0;
    }
}

// P (R x C, row-major, ld): per-ROW scale, digits written as planes S[s][row][k] with pitch cpad
// (rows >= R and k >= C are zero).  One block per (padded) row.
__global__ void __launch_bounds__(256)
oz_slice_rows_kernel(const double* __restrict__ P, int64_t ld, int R, int C, int rpad, int cpad, signed char* __restrict__ S,
                     double* __restrict__ scale) {
    __shared__ double red[32];
    const int row = blockIdx.x;
    const size_t plane = size_t(rpad) * cpad;
    signed char* out = S + size_t(row) * cpad;
    double sc = 0.0, inv = 0.0;
    if (row < R) {
        const double* p = P + int64_t(row) * ld;
        double mx = 0.0;
        for (int k = threadIdx.x * 2; k < C; k += 512) {
            mx = fmax(mx, fabs(p[k]));
            if (k + 1 < C) mx = fmax(mx, fabs(p[k + 1]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
        oz_scale_of(mx, sc, inv);
    }
    if (threadIdx.x == 0) scale[row] = sc;
    for (int k0 = threadIdx.x * 16; k0 < cpad; k0 += 256 * 16) {
        alignas(16) signed char buf[OZ_S][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = k0 + j;
            signed char q[OZ_S];
            const double v = (row < R && k < C) ? P[int64_t(row) * ld + k] : 0.0;
            oz_digits(v, inv, q);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) buf[s][j] = q[s];
        }
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) *reinterpret_cast<int4*>(out + s * plane + k0) = *reinterpret_cast<const int4*>(buf[s]);
    }
}

// column maxima of |P| over the selected columns (bit pattern of a non-negative double is monotone)
__global__ void __launch_bounds__(256)
oz_colmax_kernel(const double* __restrict__ P, int64_t ld, int R, const int* __restrict__ colmap, int ncols, int rows_per_block,
                 unsigned long long* __restrict__ mx) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= ncols) return;
    const int col = colmap ? colmap[j] : j;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(R, r0 + rows_per_block);
    double m = 0.0;
    for (int r = r0; r < r1; ++r) m = fmax(m, fabs(P[int64_t(r) * ld + col]));
    if (m > 0.0 || m != m) atomicMax(&mx[j], static_cast<unsigned long long>(__double_as_longlong(m != m ? INFINITY : m)));
}

// P (R x C, row-major, ld): per-COLUMN scale over the selected columns, digits written TRANSPOSED as
// planes S[s][j][r] with pitch rpad (j >= ncols and r >= R are zero).  64 x 64 tiles through shared memory.
__global__ void __launch_bounds__(256)
oz_slice_cols_kernel(const double* __restrict__ P, int64_t ld, int R, const int* __restrict__ colmap, int ncols, int npad, int rpad,
                     const unsigned long long* __restrict__ mx, signed char* __restrict__ S, double* __restrict__ scale) {
    constexpr int PITCH = 80;
    __shared__ __align__(16) signed char tile[OZ_S][64][PITCH];
    const int j0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int jl = threadIdx.x & 63, rp = threadIdx.x >> 6;
    const int j = j0 + jl;
    double sc = 0.0, inv = 0.0;
    int col = 0;
    if (j < ncols) {
        oz_scale_of(__longlong_as_double(static_cast<long long>(mx[j])), sc, inv);
        col = colmap ? colmap[j] : j;
    }
    if (blockIdx.y == 0 && rp == 0) scale[j] = sc;
#pragma unroll 4
    for (int rr = rp; rr < 64; rr += 4) {
        const int r = r0 + rr;
        const double v = (j < ncols && r < R) ? P[int64_t(r) * ld + col] : 0.0;
        signed char q[OZ_S];
        oz_digits(v, inv, q);
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) tile[s][jl][rr] = q[s];
    }
    __syncthreads();
    const size_t plane = size_t(npad) * rpad;
    for (int c = threadIdx.x; c < OZ_S * 64 * 4; c += 256) {
        const int s = c >> 8, jj = (c >> 2) & 63, part = c & 3;
        *reinterpret_cast<int4*>(S + s * plane + size_t(j0 + jj) * rpad + r0 + part * 16) =
            *reinterpret_cast<const int4*>(&tile[s][jj][part * 16]);
    }
}

typedef CUresult (*OzEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static OzEncodeFn oz_encode_fn() {
    static OzEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<OzEncodeFn>(p);
    }
    return fn;
}
// planes [8 * rows_pad][kpad] bytes viewed as one 2-D byte tensor; box = box_rows x 128 bytes, SWIZZLE_128B
static int oz_make_map(CUtensorMap* map, const void* base, int64_t rows_total, int64_t kpad, int box_rows) {
    OzEncodeFn fn = oz_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return 1; }
    cuuint64_t dims[2]    = {cuuint64_t(kpad), cuuint64_t(rows_total)};
    cuuint64_t strides[1] = {cuuint64_t(kpad)};
    cuuint32_t box[2]     = {cuuint32_t(OZ_BK), cuuint32_t(box_rows)};
    cuuint32_t estr[2]    = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("ozaki: cuTensorMapEncodeTiled(rows=%lld kpad=%lld) failed: %d", (long long)rows_total, (long long)kpad, int(rc)); return 1; }
    return 0;
}

}  // namespace fb200

using namespace fb200;

extern "C" int64_t fb200_ozaki_pad(int64_t n, int tile) { return round_up(n < 1 ? 1 : n, tile); }

// K splits: at least what exact int32 accumulation needs, then whatever balances the SMs best
extern "C" int fb200_ozaki_splits(int64_t Mg, int64_t Ng, int64_t K) {
    const int64_t nkb = (K + OZ_BK - 1) / OZ_BK;
    const int64_t tiles = ((Mg + OZ_BM - 1) / OZ_BM) * ((Ng + OZ_BN - 1) / OZ_BN);
    const int G = sm_count();
    const int64_t max_blocks = OZ_MAX_KSPLIT / OZ_BK;     // k blocks one split may accumulate exactly
    int smin = int((nkb + max_blocks - 1) / max_blocks);
    if (smin < 1) smin = 1;
    int best = smin;
    double best_eff = -1.0;
    for (int s = smin; s <= smin + 7 && s <= MAX_SPLIT; ++s) {
        if (s > smin && nkb / s < 16) break;
        const int64_t items = tiles * s;
        const double eff = double(items) / double(((items + G - 1) / G) * G);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
        if (eff >= 0.97) break;
    }
    return best;
}

// row-scaled digit planes of P (R x C): S is [8][rpad][cpad] int8 with rpad = pad(R, 128), cpad = pad(C, 128)
extern "C" int fb200_ozaki_slice_rows(const double* P, int64_t ld, int64_t R, int64_t C, void* S, double* scale, void* stream) {
    if (R <= 0 || C <= 0 || R > INT32_MAX / 2 || C > INT32_MAX / 2) { set_error("ozaki_slice_rows: bad shape"); return 1; }
    const int rpad = int(round_up(R, OZ_BM)), cpad = int(round_up(C, OZ_BK));
    oz_slice_rows_kernel<<<rpad, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, ld, int(R), int(C), rpad, cpad, static_cast<signed char*>(S), scale);
    return check_launch("oz_slice_rows_kernel");
}

// column-scaled, transposed digit planes of the selected columns of P (R x C): S is [8][npad][rpad] int8
// with npad = pad(ncols, tile), rpad = pad(R, 128); scratch: ncols uint64 (device)
extern "C" int fb200_ozaki_slice_cols(const double* P, int64_t ld, int64_t R, const int* colmap, int64_t ncols, int tile, void* S,
                                      double* scale, void* scratch, void* stream) {
    if (R <= 0 || ncols <= 0 || R > INT32_MAX / 2 || (tile != OZ_BN && tile != OZ_BM)) { set_error("ozaki_slice_cols: bad shape"); return 1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int npad = int(round_up(ncols, tile)), rpad = int(round_up(R, OZ_BK));
    unsigned long long* mx = static_cast<unsigned long long*>(scratch);
    if (cudaMemsetAsync(mx, 0, size_t(ncols) * sizeof(unsigned long long), st) != cudaSuccess) { set_error("ozaki_slice_cols: memset failed"); cudaGetLastError(); return 1; }
    const int gx = int((ncols + 255) / 256);
    int gy = int((4 * sm_count() + gx - 1) / gx);
    if (gy > (R + 63) / 64) gy = int((R + 63) / 64);
    if (gy < 1) gy = 1;
    const int rows_per_block = int((R + gy - 1) / gy);
    oz_colmax_kernel<<<dim3(gx, gy), 256, 0, st>>>(P, ld, int(R), colmap, int(ncols), rows_per_block, mx);
    if (check_launch("oz_colmax_kernel")) return 1;
    oz_slice_cols_kernel<<<dim3(npad / 64, rpad / 64), 256, 0, st>>>(P, ld, int(R), colmap, int(ncols), npad, rpad, mx, static_cast<signed char*>(S), scale);
    return check_launch("oz_slice_cols_kernel");
}

// C[m][colmap ? colmap[n] : n] (+ split partials) = sum_k L[m][k] R[k][n] from the digit planes
//   LS [8][mpad][kpad], lscale[mpad]   (row-scaled planes of the left operand,  mpad = pad(Mg, 128))
//   RS [8][npad][kpad], rscale[npad]   (column-scaled transposed planes of the right operand, npad = pad(Ng, 64))
extern "C" int fb200_ozaki_gemm(const void* LS, const double* lscale, int64_t Mg, const void* RS, const double* rscale, int64_t Ng,
                                int64_t K, double* C, int64_t ldc, const int* colmap, int splits, int64_t split_stride, void* stream) {
    if (Mg <= 0 || Ng <= 0 || K <= 0) { set_error("ozaki_gemm: bad shape"); return 1; }
    const int64_t kpad = round_up(K, OZ_BK), mpad = round_up(Mg, OZ_BM), npad = round_up(Ng, OZ_BN);
    const int nkb = int(kpad / OZ_BK);
    if (splits < 1) splits = 1;
    if (splits > nkb) splits = nkb;
    if (int64_t((nkb + splits - 1) / splits) * OZ_BK > OZ_MAX_KSPLIT) { set_error("ozaki_gemm: K split too long for exact int32 accumulation"); return 1; }
    static DeviceOnce attr_once;
    if (attr_once.run([] {
            cudaError_t e = cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
            if (e != cudaSuccess) { set_error("ozaki_gemm: smem attribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
            return 0;
        }))
        return 1;
    CUtensorMap mapA, mapX;
    if (oz_make_map(&mapA, LS, OZ_S * mpad, kpad, OZ_BM) || oz_make_map(&mapX, RS, OZ_S * npad, kpad, OZ_BN)) return 1;
    const int nblk_m = int(mpad / OZ_BM), nblk_n = int(npad / OZ_BN);
    const int64_t grid = int64_t(nblk_m) * nblk_n * splits;
    if (grid > INT32_MAX) { set_error("ozaki_gemm: grid too large"); return 1; }
    ozaki_gemm_kernel<<<unsigned(grid), OZ_THREADS, OZ_SMEM, static_cast<cudaStream_t>(stream)>>>(
        mapA, mapX, lscale, rscale, C, ldc, split_stride, colmap, int(Mg), int(Ng), int(mpad), int(npad), nkb, nblk_n, nblk_m, splits);
    return check_launch("ozaki_gemm_kernel");
}

/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the random stream the reference draws its Lipschitz probes
 * from: `np.random.randn` on numpy's global legacy RandomState (reference fasta/__init__.py:102-103).
 *
 * The arithmetic lives in a third-party dependency that is not under /root/reference: numpy (unpinned in the
 * reference's setup.py:3; 2.3.5 in this image), files numpy/random/src/mt19937/mt19937.c (mt19937_gen,
 * mt19937_next, mt19937_next_double) and numpy/random/src/legacy/legacy-distributions.c (legacy_gauss), which in
 * turn calls libm's log().  This file restates those published algorithms:
 *   - MT19937 block regeneration and tempering (Matsumoto & Nishimura 1998);
 *   - 53-bit double: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53;
 *   - Marsaglia polar method with the cached second deviate (returns f*x2 first, keeps f*x1);
 *   - log(): ../fasta-python_b200/csrc/glibc_log.h, the operation-by-operation restatement of glibc 2.39's log
 *     that the CUDA kernels use (the point of this file is to check THAT header on the CPU, against libm and
 *     against numpy itself, before it runs on a GPU).
 * Parity status: pinned against np.random.randn / libm log in tests/test_rng_port.py (bit for bit).
 *
 * Build: gcc -O2 -ffp-contract=off [-mfma] -shared -fPIC (oracle/build_oracle.py).  -ffp-contract=off matters:
 * only the explicit fma() calls of the header may fuse.
 */
#include <math.h>
#include <stdint.h>

#include "glibc_log.h"

static const double LOG_TAB[256] = FB200_LOG_TAB;

#define MT_N 624
#define MT_M 397

static void mt_gen(uint32_t* mt) {
    int kk;
    uint32_t y;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    }
    for (; kk < MT_N - 1; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
}

static uint32_t mt_next(uint32_t* mt, int* pos) {
    uint32_t y;
    if (*pos == MT_N) {
        mt_gen(mt);
        *pos = 0;
    }
    y = mt[(*pos)++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_double(uint32_t* mt, int* pos) {
    int32_t a = mt_next(mt, pos) >> 5, b = mt_next(mt, pos) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* n standard normals continuing the stream (key[624], pos, has_gauss, gauss) exactly as n calls of legacy_gauss */
void fb200_ref_randn(uint32_t* key, int* pos, int* has_gauss, double* gauss, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        if (*has_gauss) {
            out[i] = *gauss;
            *has_gauss = 0;
            *gauss = 0.0;
        } else {
            double f, x1, x2, r2;
            do {
                x1 = 2.0 * mt_double(key, pos) - 1.0;
                x2 = 2.0 * mt_double(key, pos) - 1.0;
                r2 = x1 * x1 + x2 * x2;
            } while (r2 >= 1.0 || r2 == 0.0);
            f = sqrt(-2.0 * fb200_glibc_log(r2, LOG_TAB) / r2);
            *gauss = f * x1;
            *has_gauss = 1;
            out[i] = f * x2;
        }
    }
}

/* the restated log next to libm's on an array: returns the number of bit mismatches */
int64_t fb200_ref_log_mismatches(const double* x, int64_t n, double* worst) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double a = fb200_glibc_log(x[i], LOG_TAB), b = log(x[i]);
        if (fb200_asuint64(a) != fb200_asuint64(b)) {
            if (bad == 0 && worst) *worst = x[i];
            ++bad;
        }
    }
    return bad;
}

double fb200_ref_log(double x) { return fb200_glibc_log(x, LOG_TAB); }

"""Developer tool: jump polynomials of MT19937 for the parallel device stream (csrc/legacy_rng.cu).

    python tools/make_mt_jump.py          # writes fasta-python_b200/fasta/mt19937_jump.npz (about 150 KB)

MT19937's word sequence y[n] satisfies, bit position by bit position, a linear recurrence over GF(2) whose
characteristic polynomial phi(t) has degree 19937 (Matsumoto & Nishimura 1998).  Hence for any J

    y[n + J] = XOR_{j : g_j = 1} y[n + j],        g(t) = t^J mod phi(t),   deg g < 19937,

(Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008, "Efficient jump ahead for F2-linear random number
generators"), so a thread block that holds 19937 + 623 consecutive words can produce the 624 words J positions
further on without generating what lies between.  This script
  1. recovers phi by Berlekamp-Massey from 2 x 19937 + 64 output bits of numpy's own MT19937,
  2. computes g for the strides the kernels use: 64 * 16^level * digit blocks of 624 words, level 0..3, digit 1..15
     (any multiple of 64 blocks below 16^4 * 64 blocks ~ 2.6e9 words is reached with at most four jumps),
  3. checks every polynomial against numpy: jumping by convolution == generating the words in between.
Polynomials are stored as 624 little-endian 32-bit words (bit j of the polynomial = bit j % 32 of word j / 32).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "fasta-python_b200", "fasta", "mt19937_jump.npz")
DEG = 19937
UNIT_BLOCKS = 64          # stride unit: 64 blocks of 624 words
LEVELS, DIGITS = 4, 16


def _generator(seed):
    """numpy's MT19937 bit generator positioned at the end of the legacy-seeded block (pos = 624)."""
    key = np.random.RandomState(seed).get_state()[1]
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 624}}
    return bg


def raw_words(seed, n):
    """n UNTEMPERED words y[0..n) that follow the seeded state (the first regenerated block onward)."""
    bg = _generator(seed)
    out = np.empty(n, dtype=np.uint32)
    for done in range(0, n, 624):
        bg.random_raw(624)                                   # consumes one whole block: the key now IS that block
        st = bg.state["state"]
        assert st["pos"] == 624
        take = min(624, n - done)
        out[done:done + take] = st["key"][:take]
    return out


def berlekamp_massey(bits):
    """Connection polynomial C (int, bit i = c_i, c_0 = 1) of the shortest LFSR generating `bits`; returns (C, L)."""
    C, B, L, m = 1, 1, 0, 1
    R = 0                                                    # bit i = s[n - i]
    for n, s in enumerate(bits):
        R = (R << 1) | int(s)
        d = bin(C & R).count("1") & 1
        if d == 0:
            m += 1
        elif 2 * L <= n:
            T = C
            C ^= B << m
            L, B, m = n + 1 - L, T, 1
        else:
            C ^= B << m
            m += 1
        R &= (1 << (DEG + 70)) - 1
    return C, L


_SPREAD = [int("".join(b + "0" for b in format(v, "08b")[::-1])[::-1], 2) for v in range(256)]      # byte -> 16 bits, zeros interleaved


def poly_square(a):
    data = a.to_bytes((a.bit_length() + 7) // 8 or 1, "little")
    out = bytearray(2 * len(data))
    for i, byte in enumerate(data):
        v = _SPREAD[byte]
        out[2 * i] = v & 0xFF
        out[2 * i + 1] = v >> 8
    return int.from_bytes(out, "little")


def poly_mod(a, phi):
    dp = phi.bit_length() - 1
    while a.bit_length() - 1 >= dp:
        a ^= phi << (a.bit_length() - 1 - dp)
    return a


def x_pow_mod(J, phi):
    """t^J mod phi by square-and-multiply (multiplying by t is a shift)."""
    r = 1
    for bit in format(J, "b"):
        r = poly_mod(poly_square(r), phi)
        if bit == "1":
            r = poly_mod(r << 1, phi)
    return r


def poly_words(g):
    return np.frombuffer(g.to_bytes(624 * 4, "little"), dtype="<u4").copy()


def jump_numpy(y_prefix, gw):
    """y[J + i], i = 0..623, from y[0 .. 19937 + 623) and the polynomial words."""
    acc = np.zeros(624, dtype=np.uint32)
    bits = np.unpackbits(gw.view(np.uint8), bitorder="little")[:DEG]
    for j in np.nonzero(bits)[0]:
        acc ^= y_prefix[j:j + 624]
    return acc


def main():
    t0 = time.time()
    # 1. phi: bit 0 of the untempered words is a linear functional of the state; its minimal polynomial is phi
    y = raw_words(4357, 2 * DEG + 700)
    C, L = berlekamp_massey((y[:2 * DEG + 64] & 1).tolist())
    assert L == DEG, f"linear complexity {L}"
    # C(t) = 1 + c_1 t + ... + c_L t^L with s[n] = XOR c_i s[n - i]; characteristic polynomial = reversal
    phi = sum(((C >> i) & 1) << (DEG - i) for i in range(DEG + 1))
    assert phi >> DEG == 1 and phi & 1 == 1
    # check: y[n + DEG] = XOR_{i < DEG, phi_i} y[n + i] on fresh words of another seed
    z = raw_words(99, DEG + 2000)
    mask = np.array([(phi >> i) & 1 for i in range(DEG)], dtype=bool)
    for n in (0, 5, 1234):
        assert np.bitwise_xor.reduce(z[n:n + DEG][mask]) == z[n + DEG], "phi does not annihilate the sequence"
    print(f"phi recovered (degree {DEG}, weight {bin(phi).count('1')}) in {time.time() - t0:.1f} s", flush=True)

    polys = np.zeros((LEVELS, DIGITS, 624), dtype=np.uint32)
    for level in range(LEVELS):
        for digit in range(1, DIGITS):
            J = 624 * UNIT_BLOCKS * (16 ** level) * digit
            g = x_pow_mod(J, phi)
            polys[level, digit] = poly_words(g)
        print(f"level {level}: strides of {UNIT_BLOCKS * 16 ** level} blocks done ({time.time() - t0:.1f} s)", flush=True)

    # 3. verification against numpy's generator: direct for the first two levels (<= 15 * 1024 blocks), by composition
    #    (jump(a) then jump(b) == jump(a + b)) for the upper levels, and one direct long jump per upper level
    seed = 77
    pre = raw_words(seed, DEG + 624)
    longest = 624 * UNIT_BLOCKS * 16 * 15 + 624
    ref = raw_words(seed, longest)
    for level in (0, 1):
        for digit in range(1, DIGITS):
            J = 624 * UNIT_BLOCKS * (16 ** level) * digit
            assert np.array_equal(jump_numpy(pre, polys[level, digit]), ref[J:J + 624]), (level, digit)
    print(f"levels 0-1 verified against direct generation ({time.time() - t0:.1f} s)", flush=True)
    for level in (2, 3):
        J = 624 * UNIT_BLOCKS * (16 ** level)
        bg = _generator(seed)
        bg.random_raw(624)                                   # y[0..623] consumed: the key is block 0 of y
        left = J
        while left > 0:                                      # advance J words; the key is then y[J .. J + 623]
            step = min(left, 1 << 26)
            bg.random_raw(step)
            left -= step
        st = bg.state["state"]
        assert st["pos"] == 624
        assert np.array_equal(jump_numpy(pre, polys[level, 1]), st["key"]), f"level {level} digit 1"
        # digits by composition: g_(d) = g_(1)^d mod phi, spot-check d = 2 and 15 through the group law
        g1 = int.from_bytes(polys[level, 1].tobytes(), "little")
        for d in (2, 7, 15):
            gd = 1
            for _ in range(d):
                # carry-less product g1 * gd mod phi via shift-and-add over the set bits of the sparser operand
                prod = 0
                a, b = g1, gd
                bits = [i for i in range(b.bit_length()) if (b >> i) & 1]
                for i in bits:
                    prod ^= a << i
                gd = poly_mod(prod, phi)
            assert np.array_equal(poly_words(gd), polys[level, d]), f"level {level} digit {d}"
        print(f"level {level} verified ({time.time() - t0:.1f} s)", flush=True)

    np.savez_compressed(OUT, polys=polys, unit_blocks=UNIT_BLOCKS, levels=LEVELS, digits=DIGITS, degree=DEG)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())

"""CPU: the arithmetic the slicing code of the int8 solver relies on (starfish_b200/csrc/sfb_internal.cuh,
oz_slice_block), restated in numpy and checked against the definition of the balanced radix-256 digits.

    q   = rint(L * 2^(47-e))                                   a 48-bit signed integer
    q   = sum_t b_t 256^t,  b_t in [-128, 127]                 (unique)
    GPU: u = low 48 bits of bits( fma(L, 2^(47-e), 1.5*2^52 + 0x808080808080) ),  b_t = int8( byte_t(u) XOR 0x80 )
"""
import numpy as np

MAGIC = 6755399441055744.0 + 141289400074368.0     # 1.5 * 2^52 + 0x808080808080


def digits_by_definition(q):
    """Balanced digits, least significant first, by repeated remainder (the round-1 code of the slicing kernel)."""
    q = q.astype(np.int64).copy()
    out = []
    for _ in range(6):
        d = ((q + 128) & 0xFF) - 128
        out.append(d.astype(np.int64))
        q = (q - d) >> 8
    assert np.all(q == 0)
    return out


def digits_by_magic(x):
    """x = L * 2^(47-e) (exact: a power-of-two scale).  One addition, bit pattern, XOR — as on the device."""
    y = x + MAGIC                                   # the fma's single rounding: x is exact, so this is rint(x) + bias
    u = y.view(np.uint64) & np.uint64(0xFFFFFFFFFFFF)
    u ^= np.uint64(0x808080808080)
    return [((u >> np.uint64(8 * t)) & np.uint64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64) for t in range(6)]


def test_magic_constant_is_exact():
    assert MAGIC == float(3 * 2 ** 51 + 0x808080808080) and 3 * 2 ** 51 + 0x808080808080 < 2 ** 53


def test_biased_magic_add_gives_the_balanced_digits():
    rng = np.random.default_rng(11)
    # |q| < 2^46 is what an accepted factorisation guarantees; include ties, tiny values, zeros, both signs
    x = np.concatenate([
        rng.uniform(-2.0 ** 46, 2.0 ** 46, 200000),
        rng.uniform(-300.0, 300.0, 50000),
        rng.integers(-2 ** 20, 2 ** 20, 50000).astype(np.float64) + 0.5,        # ties: round to even
        np.array([0.0, -0.0, 0.5, -0.5, 1.5, 127.0, 128.0, -128.0, -129.0, 32767.5, 2.0 ** 46 - 1, -(2.0 ** 46 - 1)]),
    ])
    q = np.rint(x).astype(np.int64)
    ref = digits_by_definition(q)
    got = digits_by_magic(x)
    for t in range(6):
        assert np.array_equal(ref[t], got[t]), t
    # and they recombine to q
    assert np.array_equal(sum(got[t] << (8 * t) for t in range(6)), q)


def test_representable_range():
    lim = 0x7F7F7F7F7F7F
    for q in (lim, -lim, lim - 1, -(lim - 1)):
        d = digits_by_magic(np.array([float(q)]))
        assert sum(int(d[t][0]) << (8 * t) for t in range(6)) == q

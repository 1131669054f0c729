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


# ---------------------------------------------------------------------------------------------------------------
# The MMA schedule of a 32-deep chunk (starfish_b200/csrc/ozaki.cu, oz_issue_chunk): A digit slab sa meets every B
# digit slab sb with sa + sb <= 6; a RUN of consecutive B slabs becomes ONE wide MMA into the accumulators of the
# consecutive anti-diagonals sa+sb0 ... (runs of 5 and 6 split 2+3 / 3+3; pieces of 1-2 slabs are issued in phase 0,
# of 3-4 slabs in phase 1).  Restated here and checked for every pair of digit-slab masks: each product exactly once,
# into the right accumulator, no piece wider than 4 slabs (N <= 256), nothing touching a skipped (all-zero) slab.
# ---------------------------------------------------------------------------------------------------------------
def schedule(fa, fb):
    pieces = []          # (phase, sa, sb0, len, first accumulator)
    for phase in (0, 1):
        for want in (2 * phase + 1, 2 * phase + 2):
            for sa in range(6):
                if not (fa >> sa) & 1:
                    continue
                m = fb & ((1 << (7 - sa)) - 1) & 0x3F
                for s in range(6):
                    b0 = (m >> s) & 1
                    prev = (m >> (s - 1)) & 1 if s else 0
                    if not (b0 and not prev):
                        continue
                    ln, run = 1, 1
                    for t in range(s + 1, 6):
                        run &= (m >> t) & 1
                        ln += run
                    if ln <= 4:
                        if ln == want:
                            pieces.append((phase, sa, s, ln, sa + s))
                    else:
                        if ln - 3 == want:
                            pieces.append((phase, sa, s, ln - 3, sa + s))
                        if want == 3:
                            pieces.append((phase, sa, s + ln - 3, 3, sa + s + ln - 3))
    return pieces


def test_wide_run_schedule_covers_every_product_exactly_once():
    for fa in range(64):
        for fb in range(64):
            seen = {}
            for phase, sa, sb0, ln, acc0 in schedule(fa, fb):
                assert 1 <= ln <= 4 and (ln <= 2) == (phase == 0)
                for j in range(ln):
                    sb, acc = sb0 + j, acc0 + j
                    assert (fa >> sa) & 1 and (fb >> sb) & 1, "a skipped slab is touched"
                    assert acc == sa + sb and acc <= 6
                    seen[(sa, sb)] = seen.get((sa, sb), 0) + 1
            want = {(sa, sb) for sa in range(6) for sb in range(6)
                    if (fa >> sa) & 1 and (fb >> sb) & 1 and sa + sb <= 6}
            assert set(seen) == want and all(v == 1 for v in seen.values()), (fa, fb)


def test_dense_pattern_is_nine_mmas_at_the_tensor_floor():
    p = schedule(0x3F, 0x3F)
    assert len(p) == 9 and sum(ln for _, _, _, ln, _ in p) == 26
    assert min(ln for _, _, _, ln, _ in p) >= 2          # no single-slab MMA (shared-memory-bound: 48 instead of 32 cycles)
    assert [ln for ph, _, _, ln, _ in p if ph == 1][-1] >= 3   # a chunk ends with its long MMAs

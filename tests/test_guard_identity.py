"""CPU proof-by-enumeration of the integer identity behind the two-instruction range guard of the experimental
fused update kernel (csrc/spx_fused.cu, cell_update_guard2): for the high word `hi` of a quotient and the lower
bound `qlo` that pivot_div_prepare computes (0x00100001 <= qlo <= 0x7f800001, the last value meaning "always
take the exact path"),

    qlo <= (hi & 0x7fffffff) <= 0x7f800000      (pivot_div_unchecked, csrc/spx_common.cuh)
 == (2 * hi - 2 * qlo) mod 2^32 < qr,   qr = 0 if qlo > 0x7f800000 else 0xff000001 - 2 * qlo
"""
import numpy as np


def reference_guard(hi, qlo):
    hq = hi & np.uint64(0x7FFFFFFF)
    return (hq >= qlo) & (hq <= np.uint64(0x7F800000))


def two_instruction_guard(hi, qlo):
    m = np.uint64(0xFFFFFFFF)
    q2 = (qlo + qlo) & m
    qr = np.uint64(0) if qlo > np.uint64(0x7F800000) else (np.uint64(0xFF000001) - q2) & m      # qlo is a scalar
    t = (hi * np.uint64(2) + ((np.uint64(1 << 32) - q2) & m)) & m          # mad.lo.u32 hi, 2, -q2
    return t < qr


def test_guard_identity_on_edges_and_random_words():
    rng = np.random.default_rng(3)
    edges = [0, 1, 0x000FFFFF, 0x00100000, 0x00100001, 0x00100002, 0x03600000, 0x3FF00000, 0x40000000, 0x7F7FFFFF,
             0x7F800000, 0x7F800001, 0x7FF00000, 0x7FFFFFFF]
    his = np.array(sorted({e | s for e in edges for s in (0, 0x80000000)} |
                          {(e + d) & 0xFFFFFFFF for e in edges for d in (-2, -1, 1, 2) for _ in (0,)} |
                          {0x80000000, 0xFFFFFFFF, 0xFF000000, 0xFF000001, 0xFF000002}), dtype=np.uint64)
    his = np.concatenate([his, rng.integers(0, 1 << 32, 200_000, dtype=np.uint64)])
    qlos = [0x00100001, 0x00100002, 0x00200000, 0x03600000, 0x3FF00000, 0x43500000, 0x7F700000, 0x7F800000,
            0x7F800001]
    qlos += [int(q) << 20 for q in rng.integers(2, 2040, 64)]
    for qlo in qlos:
        q = np.uint64(qlo)
        want = reference_guard(his, q)
        got = two_instruction_guard(his, q)
        assert np.array_equal(want, got), hex(qlo)
        # and around the two boundaries of this qlo, both signs
        near = np.array([(qlo + d) & 0xFFFFFFFF | s for d in range(-3, 4) for s in (0, 0x80000000)] +
                        [(0x7F800000 + d) | s for d in range(-3, 4) for s in (0, 0x80000000)], dtype=np.uint64)
        assert np.array_equal(reference_guard(near, q), two_instruction_guard(near, q)), hex(qlo)
    assert not two_instruction_guard(his, np.uint64(0x7F800001)).any()       # the sentinel never passes

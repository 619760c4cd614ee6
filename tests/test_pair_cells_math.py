"""CPU: the in-cell rank of the pair cells (csrc/query_kernels.cuh pair_below / cell_below) restated in numpy and checked
against a plain unpack-and-count on every offset, for the 4-bit and 8-bit record formats. The device code counts
"offsets below `off`" with a SWAR compare whose lane top bits are shifted to distinct positions and counted by one
popc; this pins that arithmetic (lane borrow, padding with all-ones, the shift trick) without a GPU. The kernels
themselves are checked against the oracle in tests/test_gpu_pair.py."""
import numpy as np
import pytest

U = np.uint32


def swar_below(words, off, fmt):
    """pair_below<FIRST>(r, off, fmt): words = the side's three payload words"""
    H = U(0x88888888) if fmt == 4 else U(0x80808080)
    rep = U(0x11111111) if fmt == 4 else U(0x01010101)
    t = U((int(off) * int(rep)) & 0xFFFFFFFF)
    tl = t & ~H
    acc = U(0)
    for k in range(3):
        a = U(words[k])
        d = U((int(a | H) - int(tl)) & 0xFFFFFFFF)
        lt = (~a & t) | (~(a ^ t) & ~d)
        acc |= (lt & H) >> U(k)
    return bin(int(acc)).count("1")


def pack(offsets, fmt):
    bits = 4 if fmt == 4 else 8
    per = 32 // bits
    w = [0xFFFFFFFF] * 3
    for k, o in enumerate(offsets):
        sh = (k % per) * bits
        w[k // per] = (w[k // per] & ~(((1 << bits) - 1) << sh)) | (int(o) << sh)
    return [U(x & 0xFFFFFFFF) for x in w]


@pytest.mark.parametrize("fmt,slots,width", [(4, 24, 16), (8, 12, 256)])
def test_swar_rank_equals_the_plain_count(fmt, slots, width):
    rng = np.random.default_rng(fmt)
    for trial in range(300):
        n = int(rng.integers(0, slots + 1))
        offs = np.sort(rng.integers(0, width, n))                       # ascending, duplicates allowed, the top value included
        if trial % 7 == 0 and n:
            offs[-1] = width - 1                                        # a real value equal to the padding pattern
        w = pack(offs, fmt)
        for off in list(range(0, width, 1 if fmt == 4 else 7)) + [width - 1]:
            assert swar_below(w, off, fmt) == int((offs < off).sum()), (fmt, offs, off)


def test_empty_and_full_records():
    for fmt, slots, width in ((4, 24, 16), (8, 12, 256)):
        assert all(swar_below(pack([], fmt), off, fmt) == 0 for off in range(width))
        full = pack([0] * slots, fmt)
        assert swar_below(full, 0, fmt) == 0 and swar_below(full, 1, fmt) == slots and swar_below(full, width - 1, fmt) == slots

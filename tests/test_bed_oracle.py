"""CPU: the two restatements of the reference's BED loader loop (reference test/bench.cpp:67-102) agree --
oracle/bed_oracle.py (the checker of the device tokeniser) and oracle/bed_cpu.cpp (the timed CPU
baseline of bench.py's bed_ingest line)."""
import ctypes as C
import os

import numpy as np

from oracle import bed_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_and_cpp_restatements_agree_on_well_formed_bed():
    so = os.path.join(ROOT, "oracle", "libsi_bedcpu.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libsi_bedcpu.so"])
    L = C.CDLL(so)
    L.si_bed_parse_cpu.restype = C.c_size_t
    L.si_bed_parse_cpu.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(3)
    rows = []
    for i in range(5000):
        s = int(rng.integers(0, 1_000_000)); e = s + int(rng.integers(-50, 5000))
        rows.append(f"chr{int(rng.integers(1, 6))}\t{s}\t{e}" + ("\tx\t0\t+" if i % 4 == 0 else ""))
    text = ("\n".join(rows) + "\n").encode()
    names, contig, starts, ends, lines, skipped = bed_oracle.parse_bed(text, normalize=True)
    n = len(rows)
    cs, ce, cc = (np.empty(n, np.int32) for _ in range(3))
    m = L.si_bed_parse_cpu(text, len(text), cs.ctypes.data, ce.ctypes.data, cc.ctypes.data, n)
    assert m == n == lines and skipped == 0
    assert np.array_equal(cs, starts) and np.array_equal(ce, ends) and np.array_equal(cc, contig)


def test_stoi_rules_of_the_oracle():
    text = b"a\t 12\t13x\nb\t+7\t-2\nc\tz\t1\nd\t1\t\n\ne\t2147483648\t1\nf\t00012\t0013\textra\n"
    names, contig, starts, ends, lines, skipped = bed_oracle.parse_bed(text)
    assert names == ["a", "b", "f"] and starts.tolist() == [12, 7, 12] and ends.tolist() == [13, -2, 13]
    assert lines == 7 and skipped == 4
    assert bed_oracle.parse_bed(text, normalize=True, end_shift=-1)[2:4][1].tolist() == [12, 6, 12]


def test_synthetic_bed_text_round_trips_through_the_oracle():
    """workloads.bed_text (the generator bench.py and tools/bed_bench.py feed to the device tokeniser) writes what
    it says: the oracle reads back the generating columns."""
    from superintervals_b200 import workloads as W
    text, cid, s, e = W.bed_text(3000, seed=9)
    names, contig, starts, ends, lines, skipped = bed_oracle.parse_bed(bytes(text), normalize=True, end_shift=-1)
    assert lines == 3000 and skipped == 0 and len(names) <= 24
    assert np.array_equal(starts, s.astype(np.int32)) and np.array_equal(ends, (e - 1).astype(np.int32))
    assert [names[c] for c in contig] == [f"chr{c:02d}" for c in cid]


# ---- pinned to the reference's OWN loader (Bench::load_intervals, test/bench.cpp:67-102) -------------------------------
def _chr1(text):
    names, contig, starts, ends, lines, skipped = bed_oracle.parse_bed(text, normalize=True)
    assert skipped == 0
    keep = contig == names.index("chr1")
    return starts[keep], ends[keep]


def test_oracle_equals_the_golden_output_of_the_reference_loader():
    """tests/golden/bed_ref.npz holds a seeded BED text and what the reference's loader, compiled in place
    (oracle/_ref/libsi_bedref.so, tools/make_golden_bed.py), read from it: its chr1 records, min/max-normalised."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "bed_ref.npz"))
    for text, s, e in ((g["intervals_text"].tobytes(), g["a_starts"], g["a_ends"]),
                       (g["queries_text"].tobytes(), g["q_starts"], g["q_ends"])):
        os_, oe = _chr1(text)
        assert len(s) > 500 and np.array_equal(os_, s) and np.array_equal(oe, e)


def test_oracle_equals_the_live_reference_loader_on_fresh_text():
    import pytest
    if not bed_oracle.reference_available():
        pytest.skip("oracle/_ref/libsi_bedref.so not built (no /root/reference here)")
    rng = np.random.default_rng(21)
    rows = []
    for i in range(20000):
        s = int(rng.integers(0, 250_000_000)); e = max(0, s + int(rng.integers(-500, 10_000)))
        rows.append(f"chr{int(rng.integers(1, 4))}\t{s}\t{e}" + ("\tread%d\t60\t-" % i if i % 5 == 0 else ""))
    a = ("\n".join(rows) + "\n").encode()
    b = ("\n".join(rows[::3])).encode()                      # no final newline
    (as_, ae), (bs, be) = bed_oracle.reference_load(a, b)
    for text, s, e in ((a, as_, ae), (b, bs, be)):
        os_, oe = _chr1(text)
        assert len(s) > 1000 and np.array_equal(os_, s) and np.array_equal(oe, e)

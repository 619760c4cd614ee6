"""GPU: BED ingest (siParseBed / superintervals_b200.bed.parse_bed) against the CPU restatement of the
reference's tokenising (oracle/bed_oracle.py <- reference test/bench.cpp:67-102)."""
import numpy as np
import pytest

from oracle import bed_oracle

pytestmark = pytest.mark.gpu


def _bed(rng, n, contigs, extras=True, crlf=False, final_newline=True):
    rows = []
    for i in range(n):
        c = contigs[int(rng.integers(0, len(contigs)))]
        s = int(rng.integers(0, 250_000_000))
        e = s + int(rng.integers(-5 if i % 97 == 0 else 0, 10_000))
        tail = f"\tname{i}\t{int(rng.integers(0, 1000))}\t+" if extras and i % 3 == 0 else ""
        rows.append(f"{c}\t{s}\t{e}{tail}")
    eol = "\r\n" if crlf else "\n"
    txt = eol.join(rows) + (eol if final_newline else "")
    return txt.encode()


def _same(got, want):
    names, contig, starts, ends, lines, skipped = want
    assert got.names == names
    assert got.lines == lines and got.skipped == skipped
    assert np.array_equal(got.contig, contig) and np.array_equal(got.starts, starts) and np.array_equal(got.ends, ends)


@pytest.mark.parametrize("n,crlf,final_newline", [(1, False, True), (1, False, False), (1000, False, True),
                                                  (1000, True, False), (200_000, False, True)])
@pytest.mark.parametrize("normalize,end_shift", [(False, 0), (True, -1)])
def test_bed_text_parses_like_the_reference(n, crlf, final_newline, normalize, end_shift):
    from superintervals_b200.bed import parse_bed
    rng = np.random.default_rng(n + 7 * crlf)
    text = _bed(rng, n, ["chr1", "chr2", "chrX", "chrUn_KI270742v1", "1"], crlf=crlf, final_newline=final_newline)
    _same(parse_bed(text, normalize, end_shift), bed_oracle.parse_bed(text, normalize, end_shift))


def test_headers_blanks_and_broken_lines_are_skipped_and_counted():
    from superintervals_b200.bed import parse_bed
    text = (b"track name=x description=\"y\"\n# comment\tstill\tcomment\nbrowser position chr1:1-100\n\n"
            b"chr1\t10\t20\nchr1\t  +30\t40abc\textra\nchr2\t-5\t7\nchr1\tx\t9\nchr1\t5\n\t1\t2\nchr3\t1\t99999999999\n"
            b"chr2\t2147483647\t2147483647\nchr1 10 20\nchr9\t3\t-4")
    for norm, shift in ((False, 0), (True, 0), (False, -1), (False, 1)):
        _same(parse_bed(text, norm, shift), bed_oracle.parse_bed(text, norm, shift))
    t = parse_bed(text)
    assert t.names == ["chr1", "chr2", "chr9"] and t.skipped == t.lines - len(t.starts) and len(t.starts) == 5


def test_grouping_by_contig_on_the_device_is_a_stable_partition():
    from superintervals_b200.bed import parse_bed, split_by_contig
    rng = np.random.default_rng(17)
    text = _bed(rng, 120_000, ["chr7", "chr1", "chrM", "chr22", "scaffold_123", "chrX"])
    flat = parse_bed(text, True, -1)
    grp = parse_bed(text, True, -1, group_by_contig=True)
    assert grp.names == flat.names and grp.contig_offsets is not None and grp.contig_offsets[0] == 0
    assert grp.contig_offsets[-1] == len(flat.starts) and np.all(np.diff(grp.contig_offsets) >= 0)
    order = np.argsort(flat.contig, kind="stable")
    assert np.array_equal(grp.contig, flat.contig[order])
    assert np.array_equal(grp.starts, flat.starts[order]) and np.array_equal(grp.ends, flat.ends[order])
    a, b = split_by_contig(grp), split_by_contig(flat)
    assert a.keys() == b.keys() and all(np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]) for k in a)
    one = parse_bed(b"chr5\t1\t2\nchr5\t3\t4\n", group_by_contig=True)
    assert one.contig_offsets.tolist() == [0, 2] and one.starts.tolist() == [1, 3]


def test_empty_and_newline_only_inputs():
    from superintervals_b200.bed import parse_bed
    for text in (b"", b"\n", b"\n\n\n", b"chr1"):
        _same(parse_bed(text), bed_oracle.parse_bed(text))


def test_bed_to_queries_end_to_end(tmp_path):
    """File -> device parse -> per-contig index -> counts, equal to the oracle on the same records
    (the reference's bench flow, bench.cpp:200-252, for every chrom instead of chr1 only)."""
    from oracle.pyoracle import Oracle
    from superintervals_b200 import IntervalMap
    from superintervals_b200.bed import parse_bed, split_by_contig
    rng = np.random.default_rng(5)
    contigs = ["chr1", "chr2", "chr3"]
    p_ref, p_q = tmp_path / "ref.bed", tmp_path / "q.bed"
    p_ref.write_bytes(_bed(rng, 30_000, contigs))
    p_q.write_bytes(_bed(rng, 20_000, contigs, extras=False))
    ref = split_by_contig(parse_bed(str(p_ref), True, -1, group_by_contig=True))
    qry = split_by_contig(parse_bed(str(p_q), True, -1))
    assert set(ref) == set(contigs)
    for c in contigs:
        m = IntervalMap.from_arrays(*ref[c])
        assert np.array_equal(m.count_batch_np(*qry[c]), Oracle(*ref[c]).count_batch(*qry[c]))


def test_device_parser_equals_the_reference_loaders_own_output():
    """The device tokeniser against what Bench::load_intervals (reference test/bench.cpp:67-102, compiled in place by
    oracle/Makefile) read from the same text: the committed golden pair, and -- where the compiled loader travelled --
    a fresh text."""
    import os
    from superintervals_b200.bed import parse_bed
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "bed_ref.npz"))
    cases = [(g["intervals_text"].tobytes(), g["a_starts"], g["a_ends"]), (g["queries_text"].tobytes(), g["q_starts"], g["q_ends"])]
    if bed_oracle.reference_available():
        text = _bed(np.random.default_rng(5), 50_000, ["chr1", "chr2", "chr1_KI270706v1_random"])
        (s, e), _ = bed_oracle.reference_load(text)
        cases.append((text, s, e))
    for text, s, e in cases:
        t = parse_bed(text, normalize=True)
        keep = t.contig == t.names.index("chr1")
        assert t.skipped == 0 and np.array_equal(t.starts[keep], s) and np.array_equal(t.ends[keep], e)

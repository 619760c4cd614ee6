"""examples/bed_intersect.c: the reference's example flow (examples/bed-intersect-si.rs, test/bench.cpp:200-252)
in C on the drop-in library. Compiles and links on CPU; on the GPU its per-chrom totals must equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "bed_intersect")


def _compile():
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", os.path.join(ROOT, "examples", "bed_intersect.c"), "-I" + os.path.join(ROOT, "include"),
           "-L" + os.path.join(ROOT, "superintervals_b200"), "-lsuperintervals_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "superintervals_b200"), "-o", EXE]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_c_example_compiles_against_the_c_headers():
    """A plain C99 translation unit including both public headers links against the library."""
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_c_example_totals_match_the_oracle(tmp_path):
    from oracle import bed_oracle
    from oracle.pyoracle import Oracle
    _compile()
    rng = np.random.default_rng(23)

    def bed(n, contigs):
        rows = []
        for _ in range(n):
            s = int(rng.integers(0, 3_000_000)); e = s + int(rng.integers(1, 6000))
            rows.append(f"{contigs[int(rng.integers(0, len(contigs)))]}\t{s}\t{e}")
        return ("\n".join(rows) + "\n").encode()

    ta, tq = bed(40_000, ["chr1", "chr2", "chrX"]), bed(25_000, ["chr2", "chr1", "chrY"])
    pa, pq = tmp_path / "a.bed", tmp_path / "q.bed"
    pa.write_bytes(ta); pq.write_bytes(tq)
    out = subprocess.run([EXE, str(pa), str(pq)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {l.split("\t")[0]: l.split("\t") for l in out.stdout.strip().splitlines()}
    an, ac, as_, ae, _, _ = bed_oracle.parse_bed(ta, True, -1)
    qn, qc, qs_, qe_, _, _ = bed_oracle.parse_bed(tq, True, -1)
    total = 0
    for name in an:
        if name not in qn:
            assert name not in got
            continue
        ma, mq = ac == an.index(name), qc == qn.index(name)
        want = int(Oracle(as_[ma], ae[ma]).count_batch(qs_[mq], qe_[mq]).sum())
        assert got[name][3] == f"{want} found" and got[name][4] == f"{want} counted", (name, got[name], want)
        total += want
    assert got["total"][1] == f"{total} found" and got["total"][2] == f"{total} counted"

"""GPU: the multi-device entry points of the C ABI (csrc/multi.cu) on every visible device -- one device on
the round-end box, several under `gpurun --gpus N`. Same answers as the oracle; NCCL traffic whenever more
than one device takes part; the C example drives the same entry points from plain C."""
import os
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import Oracle
from superintervals_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "multi_count")


def _compile():
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", os.path.join(ROOT, "examples", "multi_count.c"), "-I" + os.path.join(ROOT, "include"),
           "-L" + os.path.join(ROOT, "superintervals_b200"), "-lsuperintervals_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "superintervals_b200"), "-o", EXE]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_multi_count_and_search_match_the_oracle_on_every_visible_device():
    import torch
    from superintervals_b200.multi import MultiIndex
    s, e, qs, qe = W.config3(200_000, 150_001, 42, axis=12_000_000)
    orc = Oracle(s, e)
    want = orc.count_batch(qs, qe)
    off_o, res = orc.search_batch(qs, qe, want=("values",))
    for n in sorted({1, torch.cuda.device_count()}):
        m = MultiIndex(n).build(s, e)
        assert m.n_devices == n
        got = m.count_batch(qs, qe)
        assert np.array_equal(got.astype(np.uint64), want), n
        st = m.stats()
        assert (st["nccl_bytes"] > 0) == (n > 1)
        off, vals = m.search_values_batch_csr(qs, qe)
        assert np.array_equal(off, off_o) and np.array_equal(vals, res["values"]), n
        # tiny and empty batches, and a batch shorter than the device count
        for k in (0, 1, 3, 17):
            assert np.array_equal(m.count_batch(qs[:k], qe[:k]).astype(np.uint64), want[:k])


def test_c_program_drives_the_multi_device_entry_points():
    _compile()
    out = subprocess.run([EXE, "0", "200000", "500003"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout and "identical offsets and values" in out.stdout and "multi ok" in out.stdout

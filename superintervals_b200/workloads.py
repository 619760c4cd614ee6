"""Seeded synthetic workloads for the batch overlap-query path.

Every generator returns int32 numpy arrays with END-INCLUSIVE coordinates,
``(starts, ends, qs, qe)``: the stored intervals and the query ranges, in
generation (shuffled) order. ``sort_by_start`` gives the position-sorted
variant the reference benchmarks beside it
(reference test/generate_test_intervals.py:43-50 runs both).

C1  restates reference test/generate_test_intervals.py:21-41 (1M x 1M, chr1).
C2  10M read-length intervals (150 bp - 10 kb) x 100M range queries.
C3  heavy-tailed nested intervals (Pareto alpha=1.1, 50 bp - 1 Mb).
C4  24-contig whole-genome set (GRCh38 lengths), chromosome-partitioned.
C5  stabbing queries against a dense index sized to HBM.
"""
from __future__ import annotations

import numpy as np

CHR1_LEN = 250_000_000  # reference test/genome.txt:1

# GRCh38 chr1..22, X, Y (reference ships only chr1)
GRCH38 = np.array([248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
                   145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
                   101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
                   156040895, 57227415], dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def sort_by_start(starts, ends):
    """Position-sort like `bedtools sort` (chrom, start): stable on start only."""
    order = np.argsort(starts, kind="stable")
    return _i32(starts[order]), _i32(ends[order])


def _lognormal_sizes(rng, mean, variance, size):
    # reference generate_test_intervals.py:21-27: scipy lognorm(sigma, loc=mean, scale=e^mu)
    mu = np.log(mean ** 2 / np.sqrt(variance + mean ** 2))
    sigma = np.sqrt(np.log(1 + variance / mean ** 2))
    return (mean + rng.lognormal(mu, sigma, size)).astype(np.int64)


def config1(n=1_000_000, seed=0):
    """Every query overlaps >=1 stored interval; ~17 hits/query at n = 1M."""
    rng = np.random.default_rng(seed)
    lq = _lognormal_sizes(rng, 1000, 5000, n)
    lr = _lognormal_sizes(rng, 1000, 5000, n)
    pos = rng.integers(1, CHR1_LEN + 1, n)
    q_end = pos + lq                                  # half-open BED end
    lo = np.maximum(1, pos - lr - 1)
    ref_start = lo + (rng.random(n) * (q_end - 1 - lo + 1)).astype(np.int64)   # U{lo..q_end-1}
    ref_end = ref_start + lr
    # BED half-open -> inclusive (reference test/bench.cpp:210,220)
    return _i32(ref_start), _i32(ref_end - 1), _i32(pos), _i32(q_end - 1)


def _log_uniform(rng, lo, hi, size):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size)).astype(np.int64)


def config2_intervals(n=10_000_000, seed=2, axis=CHR1_LEN):
    """Read-length intervals: length log-uniform[150, 10 kb], start uniform on the axis."""
    rng = np.random.default_rng([seed, 0])
    ln = np.clip(_log_uniform(rng, 150, 10_000, n), 150, 10_000)
    s = (rng.random(n) * (axis - ln)).astype(np.int64)
    return _i32(s), _i32(s + ln - 1)


def config2_queries(nq=100_000_000, seed=2, axis=CHR1_LEN, shard=0):
    """Range queries: length log-uniform[1, 10 kb]. `shard` selects an independent stream
    (one per rank under weak scaling)."""
    rng = np.random.default_rng([seed, 1, shard])
    lq = np.clip(_log_uniform(rng, 1, 10_000, nq), 1, 10_000)
    q = (rng.random(nq) * (axis - lq)).astype(np.int64)
    return _i32(q), _i32(q + lq - 1)


def config2(n=10_000_000, nq=100_000_000, seed=2, axis=CHR1_LEN):
    """Read-length intervals x range queries on one 250 Mb axis (count only)."""
    s, e = config2_intervals(n, seed, axis)
    qs, qe = config2_queries(nq, seed, axis)
    return s, e, qs, qe


def _pareto_len(rng, size, xmin=50, alpha=1.1, cap=1_000_000):
    u = rng.random(size)
    return np.minimum(np.floor(xmin * (1.0 - u) ** (-1.0 / alpha)), cap).astype(np.int64)


def config3(n=4_000_000, nq=4_000_000, seed=42, axis=CHR1_LEN):
    """SV-like heavy-tailed nesting: stresses branch-array jumps and long hit runs."""
    rng = np.random.default_rng(seed)
    ln = _pareto_len(rng, n)
    s = (rng.random(n) * (axis - ln)).astype(np.int64)
    lq = _pareto_len(rng, nq)
    q = (rng.random(nq) * (axis - lq)).astype(np.int64)
    return _i32(s), _i32(s + ln - 1), _i32(q), _i32(q + lq - 1)


def config4_partition(n=100_000_000, nq=1_000_000_000, seed=4):
    """Per-contig (n_c, nq_c) apportioned by GRCh38 length; contigs are independent indexes."""
    w = GRCH38 / GRCH38.sum()
    nc = np.floor(w * n).astype(np.int64)
    qc = np.floor(w * nq).astype(np.int64)
    nc[0] += n - nc.sum()
    qc[0] += nq - qc.sum()
    return [(int(a), int(b), int(L), seed * 1000 + i) for i, (a, b, L) in enumerate(zip(nc, qc, GRCH38))]


def config4_contig(n_c, nq_c, length, seed):
    """One contig of C4: C2's length laws on that contig's axis."""
    return config2(n_c, nq_c, seed, axis=int(length))


def lpt_assign(costs, n_bins):
    """Longest-processing-time assignment of contigs to ranks (SURVEY 8e mode B)."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(n_bins)
    owner = np.zeros(len(costs), dtype=np.int64)
    for i in order:
        b = int(np.argmin(load))
        owner[i] = b
        load[b] += costs[i]
    return owner


def config5(n=1_000_000_000, nq=1_000_000_000, seed=5, axis=2_000_000_000):
    """Dense index on a single < 2^31 axis, stabbing queries qs == qe."""
    rng = np.random.default_rng(seed)
    ln = np.clip(_log_uniform(rng, 150, 10_000, n), 150, 10_000)
    s = (rng.random(n) * (axis - ln)).astype(np.int64)
    q = (rng.random(nq) * axis).astype(np.int64)
    return _i32(s), _i32(s + ln - 1), _i32(q), _i32(q)


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items owned by `rank` of `world` (query sharding, 8e mode A)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bed_text(n, seed=1, contigs=24):
    """n BED records 'chrNN\\tSSSSSSSSS\\tEEEEEEEEE\\n' as a uint8 array (26 bytes per line; the numbers are
    zero-padded, which std::stoi reads the same), plus the generating columns (contig number, start, end)."""
    rng = np.random.default_rng(seed)
    cid = rng.integers(1, contigs + 1, n)
    s = rng.integers(0, 249_000_000, n)
    e = s + rng.integers(1, 10_000, n)
    buf = np.empty((n, 26), np.uint8)
    buf[:, 0:3] = np.frombuffer(b"chr", np.uint8)
    buf[:, 3] = 48 + cid // 10
    buf[:, 4] = 48 + cid % 10
    buf[:, 5] = 9
    for k in range(9):
        buf[:, 6 + k] = 48 + (s // 10 ** (8 - k)) % 10
        buf[:, 16 + k] = 48 + (e // 10 ** (8 - k)) % 10
    buf[:, 15] = 9
    buf[:, 25] = 10
    return buf.reshape(-1), cid, s, e

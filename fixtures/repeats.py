"""Repeat-rich synthetic DNA (BASELINE cfg 5): uniform background + tandem arrays + interspersed
repeat families with per-copy divergence.  Seeded; test/bench fixture only."""
import numpy as np


def repeat_rich_text(n, seed=8, tandem_arrays=40, families=((300, 400, 0.05), (6000, 30, 0.10))):
    """-> (text uint8[n] over ACGT, list of (start, end) intervals that are repeat-derived)"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    text = acgt[rng.integers(0, 4, n, dtype=np.uint8)]
    regions = []

    def place(seq):
        p = int(rng.integers(0, max(1, n - len(seq))))
        m = min(len(seq), n - p)
        text[p:p + m] = seq[:m]
        regions.append((p, p + m))

    for _ in range(tandem_arrays):
        unit = acgt[rng.integers(0, 4, int(rng.integers(2, 200)))]
        copies = int(rng.integers(10, max(11, min(10_000, n // (20 * len(unit)) + 11))))
        place(np.tile(unit, copies))
    for length, copies, div in families:
        cons = acgt[rng.integers(0, 4, length)]
        copies = min(copies, max(2, n // (4 * length)))
        for _ in range(copies):
            c = cons.copy()
            d = rng.random() * div
            mut = rng.random(length) < d
            c[mut] = acgt[rng.integers(0, 4, int(mut.sum()))]
            place(c)
    return text, regions


def repeat_queries(text, regions, nq, qlen, seed=9):
    """half of the queries start inside repeat-derived regions, half anywhere"""
    rng = np.random.default_rng(seed)
    n = len(text)
    starts = np.empty(nq, dtype=np.int64)
    half = nq // 2
    reg = np.array([r for r in regions if r[1] - r[0] > qlen], dtype=np.int64)
    pick = reg[rng.integers(0, len(reg), half)]
    starts[:half] = pick[:, 0] + (rng.random(half) * (pick[:, 1] - pick[:, 0] - qlen)).astype(np.int64)
    starts[half:] = rng.integers(0, n - qlen, nq - half)
    idx = starts[:, None] + np.arange(qlen)[None, :]
    qbytes = text[idx].reshape(-1)
    qoff = np.arange(nq + 1, dtype=np.uint64) * np.uint64(qlen)
    return np.ascontiguousarray(qbytes), qoff

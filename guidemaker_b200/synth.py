"""Seeded synthetic genomes for the benchmark configurations of BASELINE.json:
i.i.d. bases with the named GC content, plus a small overlay of duplicated segments (exact-duplicate
guides, seed duplicates) and N runs (check_target rejections) that an i.i.d. genome never produces."""
from __future__ import annotations

import numpy as np

_COMP = np.arange(256, dtype=np.uint8)
_COMP[np.frombuffer(b"ACGTN", np.uint8)] = np.frombuffer(b"TGCAN", np.uint8)

CONFIGS = {
    # name: (total bases, records, gc, seed)
    "c2_bacterial_6.3Mb": (6_300_000, 1, 0.66, 2),
    "c4_yeast_12Mb": (12_000_000, 16, 0.38, 4),
    "c5_arabidopsis_120Mb": (120_000_000, 5, 0.36, 5),
    "tiny_200kb": (200_000, 2, 0.50, 1),
}


class Record:
    """Duck-typed SeqRecord: .id, str(.seq), len()."""

    def __init__(self, id: str, seq: str):
        self.id, self.seq = id, seq

    def __len__(self):
        return len(self.seq)


def synthetic_genome(total: int, records: int, gc: float, seed: int, dup_frac: float = 0.01, n_frac: float = 0.0005):
    """-> list[Record] (upper-case ASCII)"""
    rng = np.random.default_rng(seed)
    sizes = np.full(records, total // records, dtype=np.int64)
    sizes[: total - int(sizes.sum())] += 1
    out = []
    for r, n in enumerate(sizes):
        n = int(n)
        s = rng.choice(np.frombuffer(b"GCAT", np.uint8), size=n, p=[gc / 2, gc / 2, (1 - gc) / 2, (1 - gc) / 2])
        copied = 0
        while copied < dup_frac * n and n > 12000:           # duplicated 1-5 kb segments, half reverse-complemented
            ln = int(rng.integers(1000, 5001))
            src, dst = int(rng.integers(0, n - ln)), int(rng.integers(0, n - ln))
            seg = s[src:src + ln].copy()
            if rng.random() < 0.5:
                seg = _COMP[seg[::-1]]
            s[dst:dst + ln] = seg
            copied += ln
        masked = 0
        while masked < n_frac * n and n > 1000:              # N runs
            ln = int(rng.integers(1, 200))
            st = int(rng.integers(0, n - ln))
            s[st:st + ln] = ord("N")
            masked += ln
        out.append(Record("synth%d" % (r + 1), s.tobytes().decode("ascii")))
    return out


def config_genome(name: str):
    total, records, gc, seed = CONFIGS[name]
    return synthetic_genome(total, records, gc, seed)

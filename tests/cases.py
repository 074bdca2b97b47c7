"""Deterministic input cases shared by oracle/make_golden.py and the tests.

Every case is reproducible from its name alone, so the golden fixtures only
need to store the reference's OUTPUTS.
"""
from __future__ import annotations

import numpy as np

from audfprint_b200.synth import synth_track, synth_query, SR

# (name, seed, seconds)
NOISE_CASES = [("s%d_%ds" % (seed, secs), seed, secs)
               for seed in (0, 1, 2) for secs in (10, 30, 60)]


def adversarial_pcm(name: str) -> np.ndarray:
    """int16 PCM for the edge cases SURVEY.md §8c lists."""
    if name == "zeros":                      # identically zero signal
        return np.zeros(3 * SR, np.int16)
    if name == "silence_gap":                # 3 s of exact digital silence inside a track
        x = synth_track(7, 12.0).copy()
        x[4 * SR:7 * SR] = 0
        return x
    if name.startswith("short"):             # N <= n_fft: multi-bounce reflection
        n = int(name[5:])
        return synth_track(11, 1.0)[:n].copy()
    if name == "ragged":                     # N not a multiple of the hop
        return synth_track(5, 10.0)[:100003].copy()
    if name == "sine":                       # constant-amplitude sine (plateaus / exact ties)
        t = np.arange(8 * SR)
        return np.round(8000 * np.sin(2 * np.pi * 1000.0 * t / SR)).astype(np.int16)
    if name == "square":                     # clipped periodic signal
        t = np.arange(6 * SR)
        return (12000 * np.sign(np.sin(2 * np.pi * 441.0 * t / SR))).astype(np.int16)
    if name == "impulses":                   # sparse clicks over digital silence
        x = np.zeros(6 * SR, np.int16)
        x[::3001] = 20000
        return x
    raise KeyError(name)


ADVERSARIAL = ["zeros", "silence_gap", "short1", "short2", "short100", "short256", "short257",
               "short300", "short511", "short512", "short513", "ragged", "sine", "square",
               "impulses"]

DENSITY_CASES = [("s3_20s_d70", 3, 20, 70.0, 8), ("s3_20s_d100", 3, 20, 100.0, 10),
                 ("s4_20s_d7", 4, 20, 7.0, 3)]   # (name, seed, secs, density, fanout)

# small database used for the probe / match goldens
DB_NTRACKS = 40
DB_TRACK_SECONDS = 20.0
DB_QUERIES = 12          # query j is cut from track (j * 3) % DB_NTRACKS
DB_DEPTH = 20            # small depth so that some buckets overflow
DB_HASHBITS = 20
DB_MAXTIMEBITS = 14
DB2_HASHBITS = 12         # aliasing + overflowing buckets (cf. Makefile:83-85 hashbits=16)
DB2_DEPTH = 8


def db_track(i: int) -> np.ndarray:
    return synth_track(100 + i, DB_TRACK_SECONDS)


def db_query(j: int, noise_sigma: float = 0.01):
    trk = (j * 3) % DB_NTRACKS
    pcm, off = synth_query(db_track(trk), j, seconds=10.0, noise_sigma=noise_sigma)
    return pcm, trk, off

# host-side table bookkeeping sequence (oracle/make_golden_table_ops.py)
TABLE_OPS_SEED = 2718
TABLE_OPS_HASHBITS = 10
TABLE_OPS_DEPTH = 6
TABLE_OPS_ROWS = 600

# configs[3] / configs[2] geometry (oracle/make_golden_long.py): 180 s tracks, a table whose
# 12 time bits alias for them (95 s), 10 s and 200 s queries (the latter > 21k hashes at 4 shifts)
LONG_SECONDS = 180.0
LONG_SEEDS = [300, 301, 302]
LONG_HASHBITS = 16
LONG_DEPTH = 20
LONG_MAXTIMEBITS = 12
LONG_QUERIES = [(0, 10.0), (1, 10.0), (2, 10.0), (1, 170.0)]      # (track index, seconds)

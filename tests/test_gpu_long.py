"""GPU parity at the geometry of BASELINE configs[2]/[3] and on the K4 branches round 1 never
executed (VERDICT r1 'What's weak' #1, #2): 180 s files, 12 time bits that alias,
queries above the shared-memory sort capacity (unsorted probe), queries above the former
2^21-hit limit, search_depth above the fast path's capacity, empty / 1-row queries next to them.
Everything goes through the C ABI (Analyzer / HashTable / Matcher mirrors)."""
import os

import numpy as np
import pytest

from audfprint_b200 import Analyzer, HashTable, Matcher
from audfprint_b200.synth import synth_track, pcm_to_float
from oracle import afp_oracle as orc
from tests import cases
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden_long():
    return np.load(os.path.join(GOLDEN, "long.npz"))


def _table(g, depth=None):
    hashbits, d0, mtb = (int(x) for x in g["db/params"])
    table = np.zeros((1 << hashbits, d0), np.uint32)
    counts = np.zeros(1 << hashbits, np.int32)
    table[g["db/buckets"]] = g["db/rows"]
    counts[g["db/buckets"]] = g["db/counts"]
    ht = HashTable(hashbits=hashbits, depth=d0, maxtime=1 << mtb)
    ht.table, ht.counts, ht.hashesperid = table, counts, g["db/hashesperid"].astype(np.uint32)
    ht.names = ["long%d" % s for s in cases.LONG_SEEDS]
    return ht


def _same_rows(got, want, exact):
    if exact:
        return np.array_equal(got, want)
    return got.shape == want.shape and sorted(map(tuple, got)) == sorted(map(tuple, want)) and \
        np.array_equal(got[:, 1], want[:, 1])


def test_180s_files_vs_live_reference_golden(golden_long):
    """configs[3] file length (T = 7752): shifts 1 for three seeds, shifts 4 for one, in ONE batch
    per shift setting, against what the live reference produced."""
    g = golden_long
    sigs = [synth_track(s, cases.LONG_SECONDS) for s in cases.LONG_SEEDS]
    an = Analyzer()
    got = an.fingerprint_batch(sigs)
    for s, h in zip(cases.LONG_SEEDS, got):
        assert np.array_equal(h, g["t%d/wf2h_s1" % s]), s
    an4 = Analyzer()
    an4.shifts = 4
    got4 = an4.fingerprint_batch(sigs)
    assert np.array_equal(got4[0], g["t%d/wf2h_s4" % cases.LONG_SEEDS[0]])
    # the other two at 4 shifts, and a ragged 180 s-ish file, against the oracle
    for s, h in list(zip(sigs, got4))[1:]:
        assert np.array_equal(h, orc.fingerprint(pcm_to_float(s), shifts=4))
    odd = synth_track(303, 181.0)[:1984777]
    assert np.array_equal(an.fingerprint_batch([odd])[0], orc.fingerprint(pcm_to_float(odd)))


def test_aliasing_table_and_long_query_vs_live_reference_golden(golden_long):
    """maxtimebits=12 with 180 s tracks (stored times alias), reference-built table; 10 s queries
    and a 170 s query of 19,742 rows (> the 16,384-row shared-memory sort: unsorted probe path)."""
    g = golden_long
    ht = _table(g)
    m = Matcher()
    m.window, m.threshcount, m.search_depth = 2, 5, 100
    qs = [g["q%d/q" % j] for j in range(len(cases.LONG_QUERIES))]
    # empty and 1-row queries in the same batch, before and after the long one (the
    # raw-count array aliases the sorted rows and was not cleared for n <= 1)
    batch = [qs[0], np.zeros((0, 2), np.int32), qs[3], qs[1][:1], qs[3], qs[2], qs[1]]
    got = m.match_batch(ht, batch)
    for q, r in zip(batch, got):
        if len(q) <= 1:
            assert len(r) == 0
    for j, k in ((0, 0), (3, 2), (2, 5), (1, 6)):
        tie_w, tie_c = g["q%d/ties" % j]
        assert _same_rows(got[k], g["q%d/rows" % j], not tie_w and not tie_c), j
        assert got[k][0, 0] == cases.LONG_QUERIES[j][0]
    assert np.array_equal(got[4], got[2])
    # get_hits of the long query, bit for bit (query row, slot) order
    hits = ht.get_hits(qs[3])
    want = orc.get_hits(ht.table, ht.counts, ht.hashbits, ht.depth, ht.maxtimebits, qs[3])
    assert np.array_equal(hits, want)
    # 1-row query next to nothing else, every persistent CTA idle but one
    one = m.match_batch(ht, [qs[1][:1]])[0]
    assert len(one) == 0


def test_query_above_the_former_hit_limit(golden_long):
    """rows * depth >= 2^21 was AFP_ERR_UNSUPPORTED in round 1; the reference has no limit
    (hash_table.py:150-176).  A 'show' of two 180 s tracks against a depth-100 table."""
    g = golden_long
    tracks = [g["t%d/wf2h_s1" % s] for s in cases.LONG_SEEDS]
    ht = HashTable(hashbits=16, depth=100, maxtime=1 << 14)
    for s, h in zip(cases.LONG_SEEDS, tracks):
        ht.store("long%d" % s, h)
    a = g["t%d/wf2h_s4" % cases.LONG_SEEDS[0]]
    b = tracks[1].copy()
    b[:, 0] += int(a[-1, 0]) + 40
    show = np.concatenate([a, b])
    assert len(show) * ht.depth >= 1 << 21
    m = Matcher()
    m.window, m.threshcount, m.search_depth = 2, 5, 100
    got = m.match_hashes(ht, show)
    want = orc.match_hashes(ht.table, ht.counts, ht.hashbits, ht.depth, ht.maxtimebits, ht.hashesperid, show,
                            window=2, threshcount=5, search_depth=100)
    assert _same_rows(got, want, False)
    assert set(got[:2, 0].tolist()) == {0, 1}


def test_search_depth_above_the_fast_path_capacity():
    """search_depth > 1024 with more than 1024 ids above threshold: the per-candidate slow path."""
    rng = np.random.default_rng(77)
    hashbits, depth, mtb, nids = 11, 24, 10, 3000
    nb = 1 << hashbits
    table = ((rng.integers(1, nids + 1, size=(nb, depth), dtype=np.int64) << mtb)
             + rng.integers(0, 600, size=(nb, depth), dtype=np.int64)).astype(np.uint32)
    counts = rng.integers(0, depth + 6, size=nb).astype(np.int32)          # some empty, some over-full
    valid = np.arange(depth)[None, :] < np.minimum(counts, depth)[:, None]
    hpi = np.bincount((table[valid] >> mtb).astype(np.int64) - 1, minlength=nids).astype(np.uint32)
    hpi = np.maximum(hpi, 1)
    ht = HashTable(hashbits=hashbits, depth=depth, maxtime=1 << mtb)
    ht.table, ht.counts, ht.hashesperid = table, counts, hpi
    q = np.stack([rng.integers(0, 400, 2500), rng.integers(0, 1 << 20, 2500)], axis=1).astype(np.int32)
    q = q[np.lexsort((q[:, 1], q[:, 0]))]
    for thresh, sdepth in ((1, 1500), (2, 1100), (1, 1024)):
        m = Matcher()
        m.window, m.threshcount, m.search_depth = 1, thresh, sdepth
        got = m.match_batch(ht, [q, q[:300]])
        for qq, r in zip([q, q[:300]], got):
            want = orc.match_hashes(table, counts, hashbits, depth, mtb, hpi, qq, window=1, threshcount=thresh,
                                    search_depth=sdepth)
            assert _same_rows(r, want, False), (thresh, sdepth, len(qq))
        ids, raw = np.unique(orc.get_hits(table, counts, hashbits, depth, mtb, q)[:, 0], return_counts=True)
        if sdepth > 1024:
            assert np.count_nonzero(raw > thresh) > 1024      # the slow path really ran


def test_spreadpeaksinvector_vs_oracle_and_known_values():
    """Analyzer.spreadpeaksinvector (audfprint_analyze.py:153-160) as a stand-alone device call."""
    an = Analyzer()
    rng = np.random.default_rng(5)
    for n, width in ((256, 30.0), (256, 4.0), (7, 1.5), (1, 4.0), (1000, 11.0)):
        v = rng.standard_normal(n)
        if n > 10:
            v[10:13] = v[10]           # plateau: the LAST element of a plateau is the maximum
        got = an.spreadpeaksinvector(v, width)
        want = orc.spread_local_maxes(v, orc.gaussian_table(n, width))
        assert got.dtype == np.float64 and np.array_equal(got, want), (n, width)
    flat = an.spreadpeaksinvector(np.zeros(16), 4.0)
    assert np.array_equal(flat, np.zeros(16))
    assert len(an.spreadpeaksinvector(np.zeros(0))) == 0

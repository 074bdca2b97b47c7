"""HashTable.store on the device (csrc/afp_store.cu) against tables the LIVE
REFERENCE built (tests/golden/table_ops.npz, match.npz: seeded `random`, overflowing buckets)
and against the reference-pinned host store() of the mirror class."""
import os
import random

import numpy as np
import pytest

from audfprint_b200 import Analyzer, HashTable, Matcher, _lib
from audfprint_b200.synth import synth_track, synth_query
from tests import cases
from tests.conftest import GOLDEN, expand_table

pytestmark = pytest.mark.gpu


class DeviceStoreTable(HashTable):
    """The mirror class with every store() routed through the batched device insert."""

    def store(self, name, timehashpairs):
        self.store_batch([name], [timehashpairs])


def test_device_store_replays_the_reference_table_ops(golden_match, capsys):
    """The scripted store / merge / remove / slot-reuse sequence of oracle/make_golden_table_ops.py
    (2^10 x 6 table: most buckets overflow, random.seed fixed) with the stores done on the device:
    the table must equal the live reference's after every step."""
    from oracle.make_golden_table_ops_replay import run
    want = np.load(os.path.join(GOLDEN, "table_ops.npz"))
    seen = []

    def record(tag, ht):
        seen.append(tag)
        assert np.array_equal(ht.table, want[tag + "/table"]), tag
        assert np.array_equal(ht.counts, want[tag + "/counts"]), tag
        assert np.array_equal(ht.hashesperid, want[tag + "/hashesperid"]), tag
        assert ["" if n is None else n for n in ht.names] == want[tag + "/names"].tolist(), tag
    ht, r9, rlate, lines = run(DeviceStoreTable, golden_match, record)
    assert seen == ["a", "b", "merged", "removed", "reused"]
    assert np.array_equal(r9, want["retrieve_track9"]) and np.array_equal(rlate, want["retrieve_late"])
    assert ht.names[3] == "late"


@pytest.mark.parametrize("splits", [[40], [10, 20, 10], [1] * 40])
def test_store_batch_builds_the_reference_tables(golden_match, splits):
    """40 tracks into the two golden databases (db: roomy; db2: 2^12 x 8, aliasing + overflow),
    in one device batch, in three, and one by one: equal to the tables the live reference built
    with random.seed(1234), and `random` ends in the state the host replay leaves."""
    gm = golden_match
    tracks = [gm["track%d/hashes" % i] for i in range(cases.DB_NTRACKS)]
    for db in ("db", "db2"):
        table, counts, hashbits, depth, mtb, hpi = expand_table(gm, db)
        random.seed(1234)
        ht = HashTable(hashbits=hashbits, depth=depth, maxtime=1 << mtb)
        k = 0
        for n in splits:
            got = ht.store_batch(["track%d" % i for i in range(k, k + n)], tracks[k:k + n])
            assert got == [len(t) for t in tracks[k:k + n]]
            k += n
        after_device = random.random()
        assert ht._dev_newer
        assert np.array_equal(ht.counts, counts) and np.array_equal(ht.table, table), db
        assert np.array_equal(ht.hashesperid, hpi) and not ht._dev_newer
        random.seed(1234)
        host = HashTable(hashbits=hashbits, depth=depth, maxtime=1 << mtb)
        for i, t in enumerate(tracks):
            host.store("track%d" % i, t)
        assert random.random() == after_device, "the RNG replay must leave `random` where store() does"
        assert np.array_equal(host.table, table)


def test_hot_buckets_and_known_names():
    """One bucket receiving thousands of entries in a batch (CTA ranking path, beyond its
    shared-memory capacity too), names already present, ids given directly."""
    rng = np.random.default_rng(3)
    rows_a = np.stack([np.arange(9000) % 4096, np.full(9000, 77)], axis=1).astype(np.int32)       # one hash
    rows_b = np.stack([rng.integers(0, 4096, 3000), rng.integers(0, 64, 3000)], axis=1).astype(np.int32)
    rows_c = np.stack([rng.integers(0, 4096, 500), np.full(500, 77)], axis=1).astype(np.int32)
    for seed in (5, 6):
        random.seed(seed)
        dev = HashTable(hashbits=10, depth=100, maxtime=4096)
        dev.store_batch(["a", "b"], [rows_a, rows_b])
        dev.store_batch(["c", "a", 1], [rows_c, rows_b[:700], rows_a[:50]])       # "a" again, id 1 == "b"
        nxt = random.random()
        random.seed(seed)
        host = HashTable(hashbits=10, depth=100, maxtime=4096)
        for name, r in (("a", rows_a), ("b", rows_b), ("c", rows_c), ("a", rows_b[:700]), (1, rows_a[:50])):
            host.store(name, r)
        assert random.random() == nxt
        assert dev.names == host.names == ["a", "b", "c"]
        assert np.array_equal(dev.hashesperid, host.hashesperid)
        assert np.array_equal(dev.counts, host.counts) and np.array_equal(dev.table, host.table)
    with pytest.raises((ValueError, _lib.AfpError)):
        big = HashTable(hashbits=8, depth=4, maxtime=1 << 14)
        big.names = ["x"] * 300000
        big.hashesperid = np.zeros(300000, np.uint32)
        big.store_batch([299999], [rows_c])           # id does not fit in 32 - 14 bits


def test_ingest_batch_on_device_equals_per_file_store_and_serves_matches():
    sigs = [synth_track(9000 + i, 12.0 + (i % 4)) for i in range(24)]
    names = ["t%d" % i for i in range(24)]
    an = Analyzer()
    random.seed(11)
    dev = HashTable(hashbits=11, depth=12, maxtime=1 << 12)          # small: buckets overflow
    an.max_frames_per_call = 4000                                     # several device calls
    n_dev = an.ingest_batch(dev, names, sigs)
    assert dev._dev_newer
    random.seed(11)
    host = HashTable(hashbits=11, depth=12, maxtime=1 << 12)
    n_host = Analyzer().ingest_batch(host, names, sigs, on_device=False)
    assert n_dev == n_host
    # matching uses the device copy as it is: no upload, no download
    ctx = _lib.context(dev.device)
    key = ctx.table_key
    qan = Analyzer()
    qan.shifts = 4
    qs = qan.fingerprint_batch([synth_query(sigs[j], j, seconds=8.0, noise_sigma=0.01)[0] for j in (3, 17, 22)])
    m = Matcher()
    got = m.match_batch(dev, qs)
    assert ctx.table_key == key and dev._dev_newer
    want = m.match_batch(host, qs)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert [int(r[0, 0]) for r in got] == [3, 17, 22]
    _ = m.match_batch(dev, qs)                                       # re-upload after `host` took the device: still right
    assert np.array_equal(dev.table, host.table) and np.array_equal(dev.counts, host.counts)
    assert np.array_equal(dev.hashesperid, host.hashesperid) and dev.names == host.names


def test_overflow_exchange_through_the_patch_form_of_the_abi(golden_match):
    """The other form of the overflow exchange in include/afp.h - afp_table_fetch_overflow
    (bucket, count, value) + afp_table_apply_patches with ONE patch per slot chosen by the caller -
    builds the same table as store_batch (which sends only the drawn slots back)."""
    import ctypes as C
    gm = golden_match
    tracks = [gm["track%d/hashes" % i] for i in range(12)]
    hashbits, depth, mtb = 10, 6, 12
    random.seed(77)
    ref = HashTable(hashbits=hashbits, depth=depth, maxtime=1 << mtb)
    ref.store_batch(["t%d" % i for i in range(12)], tracks)
    want_table, want_counts = ref.table.copy(), ref.counts.copy()
    after = random.random()

    random.seed(77)
    ctx = _lib.context(None)
    ctx.table_key = None                                   # the context's table is ours now
    ctx.check(ctx.lib.afp_table_create(ctx.h, hashbits, depth, mtb))
    roff = np.zeros(13, np.int64)
    roff[1:] = np.cumsum([len(t) for t in tracks])
    rows = np.ascontiguousarray(np.concatenate(tracks), dtype=np.int32)
    ids = np.arange(12, dtype=np.int64)
    nov = C.c_int64(0)
    ctx.check(ctx.lib.afp_table_store_batch(ctx.h, rows.ctypes.data, 1, roff.ctypes.data_as(C.POINTER(C.c_int64)), 12,
                                            ids.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nov)))
    n = int(nov.value)
    assert n > 1000                                        # 2^10 x 6: most rows overflow
    bucket, cnt, val = np.empty(n, np.uint32), np.empty(n, np.int32), np.empty(n, np.uint32)
    ctx.check(ctx.lib.afp_table_fetch_overflow(ctx.h, bucket.ctypes.data, cnt.ctypes.data, val.ctypes.data))
    slot = np.array([random.randint(0, int(c)) for c in cnt])          # the reference's own draws
    hit = np.nonzero(slot < depth)[0]
    last = {}
    for i in hit:                                          # sequential semantics: the last write of a slot wins
        last[(int(bucket[i]), int(slot[i]))] = int(val[i])
    pb = np.array([k[0] for k in last], np.uint32)
    ps = np.array([k[1] for k in last], np.int32)
    pv = np.array(list(last.values()), np.uint32)
    ctx.check(ctx.lib.afp_table_apply_patches(ctx.h, pb.ctypes.data, ps.ctypes.data, pv.ctypes.data, len(pb)))
    table = np.empty((1 << hashbits, depth), np.uint32)
    counts = np.empty(1 << hashbits, np.int32)
    ctx.check(ctx.lib.afp_table_download(ctx.h, table.ctypes.data, counts.ctypes.data))
    assert random.random() == after
    assert np.array_equal(counts, want_counts) and np.array_equal(table, want_table)

"""Host-side pieces of round 2 that need no GPU: the C replay of CPython's random.randint
(afp_mt_randint_replay), the vectorised final sort of Matcher.match_batch, the byte layout of the
sharded-table records, bench.py's host-core detection."""
import os
import random
import sys

import numpy as np

from audfprint_b200 import _lib
from audfprint_b200 import dist as afd
from audfprint_b200.matcher import Matcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mt_replay_equals_random_randint_and_leaves_the_same_state():
    """HashTable.store draws random.randint(0, count) for every row that meets a full bucket
    (hash_table.py:127-134); store_batch replays those draws in C.  Same values, same generator
    state afterwards - across block boundaries of the Mersenne twister and for every width class."""
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for trial in range(6):
        random.seed(4000 + trial)
        for _ in range(trial * 211):
            random.random()                       # start somewhere inside a 624-word block
        counts = rng.integers(1, 6000, size=30000).astype(np.int32)
        if trial == 5:
            counts[:12] = [1, 2, 3, 4, 7, 8, 15, 16, 255, 256, 2 ** 20, 2 ** 31 - 2]
        st = random.getstate()
        state = np.array(st[1], dtype=np.uint32)
        out = np.zeros(len(counts), np.int32)
        assert lib.afp_mt_randint_replay(state.ctypes.data, counts.ctypes.data, len(counts), out.ctypes.data) == 0
        want = np.array([random.randint(0, int(c)) for c in counts], np.int64)
        assert np.array_equal(out.astype(np.int64), want), trial
        tail = [random.random() for _ in range(4)]
        random.setstate((st[0], tuple(state.tolist()), st[2]))
        assert [random.random() for _ in range(4)] == tail
    bad = np.array([-1], np.int32)
    state = np.array(random.getstate()[1], dtype=np.uint32)
    assert lib.afp_mt_randint_replay(state.ctypes.data, bad.ctypes.data, 1, np.zeros(1, np.int32).ctypes.data) != 0


def test_batch_sort_by_count_equals_the_per_query_reference_call():
    rng = np.random.default_rng(1)
    nq = 3000
    cnt = rng.integers(0, 7, nq)
    roff = np.r_[0, np.cumsum(cnt)].astype(np.int64)
    rows = rng.integers(0, 12, (int(roff[-1]), 7)).astype(np.int32)          # many tied counts
    got = Matcher._sort_by_count(rows, roff)
    for i in range(nq):
        r = rows[roff[i]:roff[i + 1]]
        assert np.array_equal(got[roff[i]:roff[i + 1]], r[(-r[:, 1]).argsort(), ])     # audfprint_match.py:335


def test_shard_record_layout_matches_the_c_side():
    """record_dtype is the NumPy view of the bytes afp_shard_pack writes (include/afp.h)."""
    lib = _lib.load()
    for sd, rcap in ((100, 16), (1, 2), (37, 64), (1500, 128)):
        dt = afd.record_dtype(sd, rcap)
        assert dt.itemsize == lib.afp_shard_record_bytes(sd, rcap) == 16 + 16 * sd + 28 * rcap
        assert dt.fields["hdr"][1] == 0 and dt.fields["w"][1] == 16 and dt.fields["id"][1] == 16 + 8 * sd
        assert dt.fields["raw"][1] == 16 + 12 * sd and dt.fields["rows"][1] == 16 + 16 * sd
    assert lib.afp_shard_record_bytes(100, 3) < 0 and lib.afp_shard_record_bytes(0, 16) < 0
    recs = [{"n_above": 3, "cand": np.array([[7, 9, 0.5], [2, 6, 0.25]]), "rows": np.arange(14, dtype=np.int32).reshape(2, 7)},
            {"n_above": 0, "cand": np.zeros((0, 3)), "rows": np.zeros((0, 7), np.int32)}]
    back = afd.unpack_shard_records(afd.pack_shard_records(recs, 5, 4))
    for a, b in zip(recs, back):
        assert a["n_above"] == b["n_above"] and np.array_equal(a["cand"], b["cand"]) and np.array_equal(a["rows"], b["rows"])


def test_host_core_detection_respects_affinity():
    sys.path.insert(0, ROOT)
    import bench
    info = bench.host_cores()
    assert 1 <= info["used"] <= info["os_cpu_count"]
    assert info["used"] <= info.get("affinity", info["os_cpu_count"])
    if "cgroup_cpus" in info:
        assert info["used"] <= max(1, int(info["cgroup_cpus"] + 0.5))

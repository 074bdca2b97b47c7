"""CPU, world_size 2 over gloo: the host side of the multi-GPU paths —
file sharding / result interleaving and the merge of per-shard match records
(audfprint_b200/dist.py).  Each rank plays one table shard, computing its records
with the ORACLE (this is a test: the product computes them in K4), then one
all_gather_object + merge must reproduce the single-table rows."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from audfprint_b200 import dist as afd
from oracle import afp_oracle as orc
from tests import cases
from tests.conftest import expand_table, GOLDEN


def shard_record(table, counts, hashbits, depth, mtb, hpi, q, lo, hi, window, thresh, sdepth):
    """What one table shard publishes for one query (ids in [lo, hi) only)."""
    hits = orc.get_hits(table, counts, hashbits, depth, mtb, q)
    hits = hits[(hits[:, 0] >= lo) & (hits[:, 0] < hi)]
    if len(hits) == 0:
        return {"n_above": 0, "cand": np.zeros((0, 3)), "rows": np.zeros((0, 7), np.int32)}
    ids, raw = np.unique(hits[:, 0], return_counts=True)
    wtd = raw / hpi[ids].astype(float)
    order = np.lexsort((-ids, -wtd))[:sdepth]
    rows = orc.offset_histogram_rows(hits, ids[order], raw[order], window, thresh)
    return {"n_above": int(np.count_nonzero(raw > thresh)),
            "cand": np.stack([ids[order], raw[order], wtd[order]], axis=1).astype(np.float64),
            "rows": rows}


def _worker(rank, world, port, db, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gm = np.load(os.path.join(GOLDEN, "match.npz"))
        table, counts, hashbits, depth, mtb, hpi = expand_table(gm, db)
        nids = len(hpi)
        keys = ["q%d_%s" % (j, tag) for j in range(cases.DB_QUERIES) for tag in ("clean", "noisy")]
        lo, hi = afd.id_range(nids, rank, world)
        ok = True
        for cfg in ("a", "b"):
            window, thresh, sdepth = (int(x) for x in gm["cfg_" + cfg])
            mine = [shard_record(table, counts, hashbits, depth, mtb, hpi, gm[k + "/q"], lo, hi, window, thresh, sdepth)
                    for k in keys]
            everyone = afd.allgather_shard_records(mine, sdepth, row_cap=128)     # the one exchange step
            assert len(everyone) == world
            assert afd.gather_objects(len(mine)) == [len(keys)] * world
            for qi, k in enumerate(keys):
                merged = afd.merge_sharded_results([everyone[r][qi] for r in range(world)], sdepth)
                merged = merged[np.argsort(-merged[:, 1], kind="stable")]
                want = orc.match_hashes(table, counts, hashbits, depth, mtb, hpi, gm[k + "/q"], window=window,
                                        threshcount=thresh, search_depth=sdepth)
                ok = ok and np.array_equal(merged, want)
        # file sharding round trip
        n = 11
        mine = [int(i) * 10 for i in afd.shard_indices(n, rank, world)]
        full = afd.interleave_shards(afd.gather_objects(mine), n)
        ok = ok and full == [i * 10 for i in range(n)]
        if rank == 0:
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("db", ["db", "db2"])
def test_sharded_match_merge_world2_gloo(db):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, db, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True


def test_shard_helpers_single_process():
    assert afd.world() == (0, 1)
    assert afd.shard_indices(10, 1, 4).tolist() == [1, 5, 9]
    assert [afd.id_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    assert afd.gather_objects([1, 2]) == [[1, 2]]
    assert afd.interleave_shards([[0, 2, 4], [1, 3]], 5) == [0, 1, 2, 3, 4]
    empty = afd.merge_sharded_results([{"n_above": 0, "cand": np.zeros((0, 3)), "rows": np.zeros((0, 7))}], 100)
    assert empty.shape == (0, 7)


@pytest.mark.parametrize("db,nshards", [("db", 3), ("db2", 2), ("db", 5)])
def test_batch_merge_equals_per_query_merge(golden_match, db, nshards):
    """merge_shard_batch (rank counting over packed records, the multi-GPU product path)
    == merge_sharded_results query by query == the single-table rows."""
    gm = golden_match
    table, counts, hashbits, depth, mtb, hpi = expand_table(gm, db)
    keys = ["q%d_%s" % (j, tag) for j in range(cases.DB_QUERIES) for tag in ("clean", "noisy")]
    for cfg in ("a", "b"):
        window, thresh, sdepth = (int(x) for x in gm["cfg_" + cfg])
        per_shard = []
        for s in range(nshards):
            lo, hi = afd.id_range(len(hpi), s, nshards)
            per_shard.append([shard_record(table, counts, hashbits, depth, mtb, hpi, gm[k + "/q"], lo, hi,
                                           window, thresh, sdepth) for k in keys])
        gathered = np.stack([afd.pack_shard_records(recs, sdepth, 128) for recs in per_shard])
        rows, off = afd.merge_shard_batch(gathered, sdepth, 128)
        assert off[-1] == len(rows) and len(off) == len(keys) + 1
        for qi, k in enumerate(keys):
            one = afd.merge_sharded_results([per_shard[s][qi] for s in range(nshards)], sdepth)
            assert np.array_equal(rows[off[qi]:off[qi + 1]], one), (db, cfg, k)
            want = orc.match_hashes(table, counts, hashbits, depth, mtb, hpi, gm[k + "/q"], window=window,
                                    threshcount=thresh, search_depth=sdepth)
            assert np.array_equal(one[np.argsort(-one[:, 1], kind="stable")], want)


def test_packed_batch_records_equal_per_query_records(golden_match):
    """pack_shard_batch (from the arrays afp_fetch_match_candidates / afp_fetch_match_rows return)
    lays records out exactly like pack_shard_records (from per-query dicts)."""
    gm = golden_match
    table, counts, hashbits, depth, mtb, hpi = expand_table(gm, "db")
    keys = ["q%d_%s" % (j, tag) for j in range(cases.DB_QUERIES) for tag in ("clean", "noisy")]
    window, thresh, sdepth = (int(x) for x in gm["cfg_a"])
    recs = [shard_record(table, counts, hashbits, depth, mtb, hpi, gm[k + "/q"], 0, len(hpi) // 2,
                         window, thresh, sdepth) for k in keys]
    nq = len(recs)
    cand = np.full((nq, sdepth, 3), 7.5)                       # junk beyond `entries`, as a reused device buffer has
    cnts = np.zeros((nq, 2), np.int32)
    rows, roff = [], np.zeros(nq + 1, np.int64)
    for i, r in enumerate(recs):
        k = len(r["cand"])
        cand[i, :k] = r["cand"]
        cnts[i] = (k, r["n_above"])
        rows.append(np.asarray(r["rows"], np.int32).reshape(-1, 7))
        roff[i + 1] = roff[i] + len(rows[-1])
    got = afd.pack_shard_batch(cand, cnts, np.concatenate(rows), roff, 128)
    assert np.array_equal(got, afd.pack_shard_records(recs, sdepth, 128))
    back = afd.unpack_shard_records(got, sdepth, 128)
    for a, b in zip(back, recs):
        assert a["n_above"] == b["n_above"] and np.array_equal(a["cand"], b["cand"])
        assert np.array_equal(a["rows"], np.asarray(b["rows"], np.int32).reshape(-1, 7))
    with pytest.raises(ValueError):
        afd.pack_shard_batch(cand, cnts, np.concatenate(rows), roff, 1)

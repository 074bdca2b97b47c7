"""One-GPU probe of the shard-mode cost: the bench table restricted to 1/8 of the ids, K4 in
publish mode + device pack + device merge of that one shard, timed; run under ncu for the profile."""
import sys, time, ctypes as C
import numpy as np
ROOT = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import multiprocessing as mp
nfiles, nq = 512, 4096
pool = mp.get_context("fork").Pool(16)
tracks = bench.make_tracks(pool, 0, nfiles, 30.0)
bench._QTRACKS = tracks
qpool = mp.get_context("fork").Pool(16)
queries = qpool.map(bench._gen_query, range(nq), chunksize=32)
pool.close(); qpool.close()
import torch
from audfprint_b200 import Analyzer, HashTable, Matcher, _lib
from audfprint_b200 import dist as afd
an = Analyzer()
hs = an.fingerprint_batch(tracks)
roff = np.zeros(nfiles + 1, np.int64); roff[1:] = np.cumsum([len(h) for h in hs])
rows = np.concatenate(hs)
table, counts, hashbits, depth, mtb, hpi, ids = bench.build_big_table(rows, roff, 1000000)
ht = HashTable(hashbits=hashbits, depth=1, maxtime=1 << mtb)
ht.table, ht.counts, ht.hashesperid, ht.depth = table, counts, hpi, depth
qan = Analyzer(); qan.shifts = 4
qh = qan.fingerprint_batch([q[0] for q in queries])
qoff = np.zeros(nq + 1, np.int64); qoff[1:] = np.cumsum([len(h) for h in qh])
qrows = np.ascontiguousarray(np.concatenate(qh))
m = Matcher(); m.window = 2
dq = torch.from_numpy(qrows).cuda()
ctx = ht._sync_device()
def timeit(f, n=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
p = m._params(); tot = C.c_int64(0); offp = qoff.ctypes.data_as(C.POINTER(C.c_int64))
t_single = timeit(lambda: ctx.check(ctx.lib.afp_match_batch(ctx.h, dq.data_ptr(), 0, nq, offp, C.byref(p), C.byref(tot))))
print("single table: %.2f ms  %.0f q/s" % (t_single * 1e3, nq / t_single))
for ns in (2, 8):
    ht.restrict_device_ids(*afd.id_range(1000000, 0, ns))
    pp = m._params(); pp.publish_candidates = 1
    t_k4 = timeit(lambda: ctx.check(ctx.lib.afp_match_batch(ctx.h, dq.data_ptr(), 0, nq, offp, C.byref(pp), C.byref(tot))))
    st = Matcher.last_status(ht, nq)
    t_all = timeit(lambda: afd.match_sharded_batch(m, ht, (dq, qoff), row_cap=16, fetch=False))
    print("1/%d shard: K4 publish %.2f ms (%.0f q/s), + pack + merge %.2f ms; handover %d, members %.0f, extras %.0f"
          % (ns, t_k4 * 1e3, nq / t_k4, t_all * 1e3, int((st[:, 0] > 0).sum()), st[:, 1].mean(), st[:, 3].mean()))

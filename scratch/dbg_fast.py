import sys, types, time, numpy as np
sys.path.insert(0, '/root/repo')
src = open('/root/repo/tests/test_gpu_match_fast.py').read().replace('pytestmark = pytest.mark.gpu', '')
mod = types.ModuleType('t'); exec(compile(src, 't', 'exec'), mod.__dict__)
from audfprint_b200 import Matcher
table, counts = mod.make_table(1, 19, 100, 1 << 20, 12)
rng = np.random.default_rng(2)
qs = [mod.make_query(100 + i, 650 + 30 * i, 19) for i in range(8)]
for i, q in enumerate(qs):
    mod.plant(table, counts, q, 19, 100, 12, 5000 + i, 300 + 7 * i, 120, rng)
hpi = mod.hpi_of(table, counts, 100, 12, 1 << 20)
ht = mod.as_ht(table, counts, 19, 100, 12, hpi)
m = Matcher(); m.window, m.threshcount, m.search_depth = 2, 5, 100
for batch in (qs, qs[::-1], [qs[3]], [qs[0][:900], qs[1][:1000], qs[1][:1100], qs[1][:1150], qs[1]]):
    r = m.match_batch(ht, batch, sort=False)
    print([len(q) for q in batch])
    print(Matcher.last_status(ht, len(batch)).tolist())

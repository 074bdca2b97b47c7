"""The scripted sequence of table operations behind tests/golden/table_ops.npz.
TEST INFRASTRUCTURE: run once on the live reference class (oracle/make_golden_table_ops.py)
and replayed on the mirror class by tests/test_abi_cpu.py; imports neither."""
import random

import numpy as np

from tests import cases


def run(HashTable, gm, record):
    """The scripted sequence of table operations; `record(tag, ht)` snapshots the state.
    HashTable is the class under test (reference or mirror)."""
    random.seed(cases.TABLE_OPS_SEED)
    np.random.seed(cases.TABLE_OPS_SEED)
    geom = dict(hashbits=cases.TABLE_OPS_HASHBITS, depth=cases.TABLE_OPS_DEPTH, maxtime=1 << 12)
    a, b = HashTable(**geom), HashTable(**geom)
    for i in range(8):
        a.store("track%d" % i, gm["track%d/hashes" % i][:cases.TABLE_OPS_ROWS])
    for i in range(8, 14):
        b.store("track%d" % i, gm["track%d/hashes" % i][:cases.TABLE_OPS_ROWS])
    record("a", a)
    record("b", b)
    a.merge(b)
    record("merged", a)
    a.remove("track3")
    record("removed", a)
    a.store("late", gm["track14/hashes"][:cases.TABLE_OPS_ROWS])       # takes the freed id 3
    record("reused", a)
    lines = []
    a.list(lines.append)
    return a, a.retrieve("track9"), a.retrieve("late"), lines



"""Host-side behaviour of the API mirror that needs no GPU: report strings, pickling
hygiene, device-copy bookkeeping."""
import gzip
import os
import pickle
import subprocess
import sys

import numpy as np

from audfprint_b200 import Analyzer, HashTable, Matcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

# What the live reference's Matcher.file_match_to_msgs returns (audfprint_match.py:381-420) when
# match_file yields these rows; generated in the build container by patching match_file on the
# reference class (the strings are its output, not code).
ROWS = np.array([[1, 123, -45, 200, 0, 17, 402], [2, 9, 1033, 40, 3, 0, 12]], np.int32)
NAMES = ["a.mp3", "dir/b song.wav", "c"]
QRY = "query file.mp3"
GOLDEN_MSGS = {
    (0, 0, 1): ['query file.mp3\tdir/b song.wav', 'query file.mp3\tc'],
    (0, 0, 0): ['query file.mp3\t'],
    (0, 1, 1): ['query file.mp3\tdir/b song.wav', 'query file.mp3\tc'],
    (0, 1, 0): ['query file.mp3\t'],
    (1, 0, 1): ['Matched query file.mp3 12.3 sec 567 raw hashes as dir/b song.wav at   -1.0 s with   123 of   200 '
                'common hashes at rank  0',
                'Matched query file.mp3 12.3 sec 567 raw hashes as c at   24.0 s with     9 of    40 common hashes '
                'at rank  3'],
    (1, 0, 0): ['NOMATCH query file.mp3 12.3 sec 567 raw hashes'],
    (1, 1, 1): ['Matched    8.9 s starting at    0.4 s in query file.mp3 to time   -0.7 s in dir/b song.wav with   '
                '123 of   200 common hashes at rank  0',
                'Matched    0.3 s starting at    0.0 s in query file.mp3 to time   24.0 s in c with     9 of    40 '
                'common hashes at rank  3'],
    (1, 1, 0): ['NOMATCH query file.mp3 12.3 sec 567 raw hashes'],
}


class _Names:
    names = NAMES


def test_report_lines_equal_the_reference_including_time_range_branch():
    for (verbose, ftr, hit), want in GOLDEN_MSGS.items():
        m = Matcher()
        m.verbose, m.find_time_range = bool(verbose), bool(ftr)
        rows = ROWS if hit else ROWS[:0]
        m.match_file = lambda an, ht, q, number=None, rows=rows: (rows, 12.34, 567)
        assert m.file_match_to_msgs(Analyzer(), _Names(), QRY) == want, (verbose, ftr, hit)


def test_save_leaves_no_stand_in_module_behind(tmp_path):
    """HashTable.save registered a stub `hash_table` module and never removed it."""
    code = r"""
import sys, os
sys.path.insert(0, %r)
from audfprint_b200 import HashTable
ht = HashTable(hashbits=8, depth=4, maxtime=256)
ht.store("x", [(1, 2), (3, 4)])
assert "hash_table" not in sys.modules
ht.save(%r)
assert "hash_table" not in sys.modules, "stand-in module leaked"
if os.path.isdir(%r):
    sys.path.insert(0, %r)
    import hash_table                      # the REAL reference module, usable after our save
    ref = hash_table.HashTable(%r)
    assert ref.names == ["x"] and int(ref.counts.sum()) == 2
    ht.save(%r)                            # with the real module loaded the real class is used
    assert sys.modules["hash_table"] is hash_table
print("ok")
""" % (ROOT, str(tmp_path / "a.pklz"), REF, REF, str(tmp_path / "a.pklz"), str(tmp_path / "b.pklz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
    with gzip.open(tmp_path / "a.pklz", "rb") as f:
        assert b"hash_table" in f.read(4096)


def test_device_stamp_is_unique_per_table_and_tracks_rebinding():
    """the freshness stamp must not be reusable by another table, and rebinding the
    public arrays must change it."""
    stamps = set()
    for _ in range(50):
        ht = HashTable(hashbits=6, depth=2, maxtime=64)
        stamps.add(ht._stamp())
        del ht
    assert len(stamps) == 50
    ht = HashTable(hashbits=6, depth=2, maxtime=64)
    s0 = ht._stamp()
    ht.table = np.zeros((64, 2), np.uint32)
    s1 = ht._stamp()
    ht.counts = np.zeros(64, np.int32)
    s2 = ht._stamp()
    ht.depth = 2
    ht.store("a", [(0, 1)])
    s3 = ht._stamp()
    ht.table[0, 0] = 7
    ht.touch()
    s4 = ht._stamp()
    assert len({s0, s1, s2, s3, s4}) == 5
    clone = pickle.loads(pickle.dumps(ht))
    assert clone._stamp()[0] != ht._stamp()[0]
    assert np.array_equal(clone.table, ht.table) and clone.names == ht.names
    assert "table" in ht.__getstate__() and "_table" not in ht.__getstate__()

"""CPU, build container only: the oracle against the LIVE reference on seeds that are not in
tests/golden (skipped where /root/reference does not exist, i.e. on the GPU box).  The reference
is imported in a subprocess so that its module names (hash_table, ...) never enter this process."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from audfprint_b200.synth import synth_track, synth_query, pcm_to_float
from oracle import afp_oracle as orc

REF = os.environ.get("AFP_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "audfprint_analyze.py")),
                                reason="live reference not present")

_DRIVER = r'''
import json, random, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
import audfprint_analyze as an, audfprint_match as ma, audio_read as ar, hash_table as htm
from audfprint_b200.synth import synth_track, synth_query, pcm_to_float
pcm = {}
ar.audio_read = lambda fn, sr=None, channels=None: (pcm_to_float(pcm[fn]), 11025)
out = {}
tracks = []
for seed in %(seeds)r:
    pcm["t"] = synth_track(seed, 14.0 + seed %% 5)
    for shifts in (1, 4):
        a = an.Analyzer(); a.shifts = shifts
        out["h_%%d_%%d" %% (seed, shifts)] = np.asarray(a.wavfile2hashes("t")).tolist()
    tracks.append(np.asarray(out["h_%%d_1" %% seed], np.int32))
random.seed(4)
ht = htm.HashTable(hashbits=14, depth=6, maxtime=1 << 12)
for i, h in enumerate(tracks):
    ht.store("s%%d" %% i, h)
m = ma.Matcher(); m.window = 2; m.threshcount = 3; m.search_depth = 4
for j, seed in enumerate(%(seeds)r):
    q, _ = synth_query(synth_track(seed, 14.0 + seed %% 5), 77 + j, seconds=6.0, noise_sigma=0.01)
    pcm["q"] = q
    a = an.Analyzer(); a.shifts = 4
    qh = np.asarray(a.wavfile2hashes("q"), np.int32)
    out["q_%%d" %% seed] = qh.tolist()
    out["hits_%%d" %% seed] = ht.get_hits(qh).tolist()
    out["rows_%%d" %% seed] = m.match_hashes(ht, qh).tolist()
out["table"] = ht.table.tolist(); out["counts"] = ht.counts.tolist(); out["hpi"] = np.asarray(ht.hashesperid).tolist()
print("JSON" + json.dumps(out))
'''


def test_oracle_equals_live_reference_on_fresh_seeds():
    seeds = [5101, 5102, 5103, 5104]
    code = _DRIVER % {"root": ROOT, "ref": REF, "seeds": seeds}
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    ref = json.loads([ln for ln in run.stdout.splitlines() if ln.startswith("JSON")][0][4:])
    tracks = []
    for seed in seeds:
        d = pcm_to_float(synth_track(seed, 14.0 + seed % 5))
        for shifts in (1, 4):
            got = orc.fingerprint(d, shifts=shifts)
            assert np.array_equal(got, np.array(ref["h_%d_%d" % (seed, shifts)], np.int32).reshape(-1, 2)), (seed, shifts)
        tracks.append(orc.fingerprint(d, shifts=1))
    import random
    rng = random.Random(4)
    t = orc.Table(hashbits=14, depth=6, maxtimebits=12)
    for i, h in enumerate(tracks):
        t.store("s%d" % i, h, rng)
    assert np.array_equal(t.table, np.array(ref["table"], np.uint32)) and np.array_equal(t.counts, ref["counts"])
    hpi = np.array(ref["hpi"])
    for j, seed in enumerate(seeds):
        q, _ = synth_query(synth_track(seed, 14.0 + seed % 5), 77 + j, seconds=6.0, noise_sigma=0.01)
        qh = orc.fingerprint(pcm_to_float(q), shifts=4)
        assert np.array_equal(qh, np.array(ref["q_%d" % seed], np.int32).reshape(-1, 2))
        hits = orc.get_hits(t.table, t.counts, 14, 6, 12, qh)
        assert np.array_equal(hits, np.array(ref["hits_%d" % seed], np.int32).reshape(-1, 4))
        rows = orc.match_hashes(t.table, t.counts, 14, 6, 12, hpi, qh, window=2, threshcount=3, search_depth=4)
        want = np.array(ref["rows_%d" % seed], np.int32).reshape(-1, 7)
        assert rows.shape == want.shape and np.array_equal(rows[:1], want[:1]), seed   # best match identical
        assert sorted(map(tuple, rows[:, :4])) == sorted(map(tuple, want[:, :4]))        # same alignments

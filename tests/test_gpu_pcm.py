"""The device PCM front-end (csrc/afp_pcm.cu): down-mix + polyphase resampling,
pinned by TOLERANCE against scipy.signal.resample_poly (the reference delegates this step to
ffmpeg, audio_read.py:196-203, whose resampler no other implementation reproduces bit for bit)."""
import wave

import numpy as np
import pytest
from scipy.signal import resample_poly

from audfprint_b200 import Analyzer
from audfprint_b200 import analyzer as an_mod
from audfprint_b200.synth import synth_track

pytestmark = pytest.mark.gpu


def _host_statement(raw, nch, fs, sr):
    x = raw.reshape(-1, nch).astype(np.float32) * np.float32(1.0 / 32768.0)
    mono = x.mean(axis=1).astype(np.float32) if nch > 1 else x[:, 0]
    if sr is None or sr == fs:
        return mono
    from math import gcd
    g = gcd(sr, fs)
    return resample_poly(mono.astype(np.float64), sr // g, fs // g).astype(np.float32)


@pytest.mark.parametrize("fs,nch", [(44100, 2), (22050, 1), (48000, 2), (8000, 1), (11025, 3), (16000, 6)])
def test_frontend_matches_resample_poly(fs, nch):
    rng = np.random.default_rng(fs + nch)
    n = int(fs * 1.7) + 13
    raw = (rng.standard_normal((n, nch)) * 6000).clip(-32768, 32767).astype(np.int16)
    t = np.arange(n) / fs
    raw[:, 0] += (9000 * np.sin(2 * np.pi * 997.0 * t)).astype(np.int16)
    got = an_mod.pcm_frontend(raw, nch, fs, 11025)
    want = _host_statement(raw, nch, fs, 11025)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.max(np.abs(got - want)) < 2e-6          # float32 rounding of a float64 dot product
    dev = an_mod.pcm_frontend(raw, nch, fs, 11025, to_host=False)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)
    assert len(an_mod.pcm_frontend(raw[:0], nch, fs, 11025)) == 0


def test_wav_reader_and_file_level_api(tmp_path):
    """A stereo 44.1 kHz WAV of a known track: read (device down-mix + resample), fingerprint,
    and compare with fingerprinting the host statement of the same front-end."""
    fs = 44100
    mono11 = synth_track(4242, 8.0)
    up = resample_poly(mono11.astype(np.float64), 4, 1)
    stereo = np.stack([up * 0.9, up * 0.7], axis=1).clip(-32768, 32767).astype(np.int16)
    fn = str(tmp_path / "s.wav")
    with wave.open(fn, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(fs)
        w.writeframes(stereo.tobytes())
    d, sr = an_mod._wav_reader(fn, sr=11025, channels=1)
    want = _host_statement(stereo.reshape(-1), 2, fs, 11025)
    assert sr == 11025 and d.dtype == np.float32 and len(d) == len(want)
    assert np.max(np.abs(d - want)) < 2e-6
    d2, sr2 = an_mod._wav_reader(fn)                      # no target rate: native rate, still mono
    assert sr2 == fs and len(d2) == len(stereo)
    an = Analyzer()
    h = an.wavfile2hashes(fn)
    h_want = an.fingerprint_batch([want])[0]
    # identical samples up to 2e-6 -> practically the same landmarks (not a bit-parity claim)
    a, b = set(map(tuple, h.tolist())), set(map(tuple, h_want.tolist()))
    assert len(a & b) / max(1, len(a | b)) > 0.97 and abs(an.soundfiledur - 8.0) < 0.01

"""Seeded synthetic WSJ0-2mix-shaped mixtures.  Used by tests and bench.

No audio ships with the reference, so every parity / timing case is built from
"speech-like" noise: white Gaussian noise through a random two-pole resonator bank, gated by
a ~4 Hz syllabic envelope, then preprocessed exactly like the reference does per source
(-mean, /max|.|, zero-pad to MAX_LEN, gain 10^(dB/20); TDAA_beta/predata_fromList.py:140-177)
and summed into the mixture.
"""
import numpy as np
from scipy.signal import lfilter

from .stft_ref import preprocess_source


def speech_like(rng, n, sr=8000):
    x = rng.standard_normal(n)
    y = np.zeros(n)
    for _ in range(3):
        f = rng.uniform(200.0, 3400.0)
        r = rng.uniform(0.90, 0.98)
        a = [1.0, -2.0 * r * np.cos(2 * np.pi * f / sr), r * r]
        y += lfilter([1.0], a, x) * rng.uniform(0.3, 1.0)
    t = np.arange(n) / float(sr)
    env = 0.5 * (1.0 + np.sign(np.sin(2 * np.pi * rng.uniform(2.5, 5.0) * t + rng.uniform(0, 6.28))))
    env = lfilter([0.02], [1.0, -0.98], env)
    return y * (0.05 + env)


def make_batch(B, L, S, seed=1, num_spk=101, active_len=None, sr=8000):
    """-> dict(mix_wav [B,L] f64, sources [B,S,L] f64, spk_idx [B,S] int64 sorted, gains_db [B,S])."""
    rng = np.random.RandomState(seed)
    mix = np.zeros((B, L))
    srcs = np.zeros((B, S, L))
    idx = np.zeros((B, S), dtype=np.int64)
    gains = np.zeros((B, S))
    for b in range(B):
        idx[b] = np.sort(rng.choice(num_spk, S, replace=False))
        for s in range(S):
            n = L if active_len is None else int(rng.randint(active_len[0], active_len[1] + 1))
            n = min(n, L)
            g = rng.uniform(-2.5, 2.5)
            srcs[b, s] = preprocess_source(speech_like(rng, n, sr), L, g)
            gains[b, s] = g
        mix[b] = srcs[b].sum(0)
    return {'mix_wav': mix, 'sources': srcs, 'spk_idx': idx, 'gains_db': gains}

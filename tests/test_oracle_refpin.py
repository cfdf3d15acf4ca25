"""CPU tests that pin oracle/stft_ref.py to outputs of THE REFERENCE'S OWN CODE.

tests/golden/ref_stft_*.npz were written by tests/golden/make_ref_fixtures.py, which exec()s the pure-numpy
sqrt_hann / stft / istft of /root/reference/Cocktail/software/DL4SS_Keras/test_stft_istft.py:9-63 (the only
value-producing transform code in the reference tree; librosa itself is not vendored).  Two links:
  1. oracle (center=False, the restatement of those functions) == fixtures to 1e-6;
  2. oracle (center=True, librosa semantics: the form the product path implements) == fixtures on every frame /
     sample the two conventions share (a centred frame t starts at t*hop - n_fft/2, so reference frame i is centred
     frame i + overlap/2; the overlap-added interior is identical, only the trim differs).
"""
import os

import numpy as np
import pytest

from oracle import stft_ref as sr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['ref_stft_256_2', 'ref_stft_256_4', 'ref_stft_1024_2']


def gold(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)


def peak_err(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize('name', CASES)
def test_uncentred_oracle_reproduces_reference_functions(name):
    g = gold(name)
    n_fft, ov = int(g['n_fft']), int(g['overlap'])
    hop = n_fft // ov
    assert np.array_equal(sr.get_window('sqrt_hanning', n_fft), g['window'])
    X = sr.stft_ref(g['x'], n_fft, hop, 'sqrt_hanning', center=False)
    assert X.T.shape == g['X'].shape
    assert peak_err(X.T, g['X']) < 1e-6
    assert peak_err(sr.stft_ref(g['x'], n_fft, hop, 'sine', center=False).T, g['Xs']) < 1e-6
    assert peak_err(sr.istft_ref(g['X'].T, hop, 'sqrt_hanning', center=False), g['y']) < 1e-6
    assert peak_err(sr.istft_ref((g['mask'] * g['X']).T, hop, 'sqrt_hanning', center=False), g['ym']) < 1e-6
    assert peak_err(sr.istft_ref(g['Xs'].T, hop, 'sine', center=False), g['ys']) < 1e-6


@pytest.mark.parametrize('name', CASES)
def test_centred_oracle_agrees_with_reference_on_shared_frames(name):
    g = gold(name)
    n_fft, ov = int(g['n_fft']), int(g['overlap'])
    hop = n_fft // ov
    n = g['X'].shape[0]
    for window, key in (('sqrt_hanning', 'X'), ('sine', 'Xs')):
        S = sr.stft_ref(g['x'], n_fft, hop, window).T                    # centred, complex64
        shared = S[ov // 2: ov // 2 + n]
        assert peak_err(shared, g[key]) < 1e-6                            # complex64 rounding only
    # inverse: the reference overlap-adds frames 0 .. n-ov-1 into [0, hop*(n-1)); librosa semantics trims n_fft/2
    for window, spec, key in (('sqrt_hanning', g['X'], 'y'), ('sqrt_hanning', g['mask'] * g['X'], 'ym'),
                              ('sine', g['Xs'], 'ys')):
        y = sr.istft_ref(spec[:n - ov].T, hop, window)                   # float32 [hop*(n-ov-1)]
        ref = g[key][n_fft // 2: hop * (n - 1) - n_fft // 2]
        assert y.shape == ref.shape
        assert peak_err(y, ref) < 1e-6


def test_reference_roundtrip_identity():
    """The fixtures themselves: istft(stft(x)) returns x wherever every overlapping frame was kept."""
    g = gold('ref_stft_256_2')
    n = g['X'].shape[0]
    lo, hi = 128, 128 * (n - 2)
    assert np.abs(g['y'][lo:hi] - g['x'][lo:hi]).max() < 1e-12

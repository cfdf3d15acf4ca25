#!/usr/bin/env python
"""Fixtures produced BY THE REFERENCE'S OWN CODE (run in the build container, where /root/reference exists):

    python tests/golden/make_ref_fixtures.py

The only value-producing signal-processing code the reference ships that runs unchanged under Python 3 is the
pure-numpy `sqrt_hann` / `stft` / `istft` of Cocktail/software/DL4SS_Keras/test_stft_istft.py:9-63.  This script
exec()s exactly those lines of the file where it lies (nothing is copied into the repo), feeds them seeded synthetic
waveforms and stores inputs and outputs:

    ref_stft_256_2.npz   n_fft 256, overlap 2 (hop 128: every TDAA_beta / Torch_multi config)
    ref_stft_256_4.npz   n_fft 256, overlap 4 (hop 64: BASELINE configs[0])
    ref_stft_1024_2.npz  n_fft 1024, overlap 2 (the Keras tree's setting, test_stft_istft.py:87-89)
keys: x [L] f64, X [n,F] c128 = stft(x), mask [n,F] f32, y [n*hop] f64 = istft(X), ym = istft(mask * X), plus the
sine-window variant the same file uses (`windows`, :90): Xs = stft(x, window=sine), ys = istft(Xs, window=sine).

tests/test_oracle.py checks oracle/stft_ref.py (center=False, window 'sqrt_hanning') against them; the GPU suite
checks K1 / K6 against them (tests/test_gpu_golden.py).  /root/reference is not needed to RUN any test.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/Cocktail/software/DL4SS_Keras/test_stft_istft.py'
LINES = (9, 63)         # def sqrt_hann ... end of istft


def load_reference_functions():
    with open(REF, encoding='utf-8') as f:
        src = f.readlines()
    code = ''.join(src[LINES[0] - 1:LINES[1]])
    ns = {'np': np}
    exec(compile(code, REF, 'exec'), ns)
    return ns['sqrt_hann'], ns['stft'], ns['istft']


def speechy(rng, n):
    t = np.arange(n) / 8000.0
    x = rng.standard_normal(n) * (0.2 + 0.8 * (np.sin(2 * np.pi * 3.1 * t) > 0))
    x += 0.5 * np.sin(2 * np.pi * 440.0 * t + rng.uniform(0, 6.28)) + 0.3 * np.sin(2 * np.pi * 1730.0 * t)
    return x / np.abs(x).max()


def case(name, n_fft, overlap, L, seed, funcs):
    sqrt_hann, stft, istft = funcs
    rng = np.random.RandomState(seed)
    x = speechy(rng, L)
    X = stft(x, n_fft, overlap)
    mask = rng.uniform(0, 1, X.shape).astype(np.float32)
    y = istft(X, overlap)
    ym = istft(mask * X, overlap)
    sine = [np.sin(i * np.pi / n_fft) for i in range(n_fft)]            # test_stft_istft.py:90
    Xs = stft(x, n_fft, overlap, window=sine)
    ys = istft(Xs, overlap, window=sine)
    assert X.shape[1] == n_fft // 2 + 1 and y.shape[0] == X.shape[0] * (n_fft // overlap)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), x=x, X=X, mask=mask, y=y, ym=ym, Xs=Xs, ys=ys,
                        n_fft=n_fft, overlap=overlap, window=sqrt_hann(n_fft))


def main():
    funcs = load_reference_functions()
    case('ref_stft_256_2', 256, 2, 6000, 11, funcs)
    case('ref_stft_256_4', 256, 4, 4000, 12, funcs)
    case('ref_stft_1024_2', 1024, 2, 9000, 13, funcs)
    for f in sorted(os.listdir(HERE)):
        if f.startswith('ref_') and f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    sys.exit(main())

#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ (run from the repo root, CPU only):

    python tests/golden/make_golden.py

The reference (Python 2 + librosa + mir_eval) cannot be imported in this image, and
its tree holds no value fixtures, so the vectors come from the two independent
implementations that ARE here:
  * stft_*.npz  : torch.stft / torch.istft with the reference's settings (periodic Hann, centre,
                  reflect, onesided) -- an implementation independent of oracle/stft_ref.py; the CPU
                  suite checks the oracle against them, the GPU suite checks the CUDA kernels.
  * model_*.npz : the reference's own arithmetic (torch.nn.LSTM/GRU/Linear/Embedding, baddbmm,
                  sigmoid/tanh, MSELoss on CPU fp32) evaluated through oracle/modules_ref.py with
                  seeded weights; the weights are re-created from the seed, only inputs/outputs
                  are stored (float32, a few hundred KB in total).
Shape constants embedded in the reference (39936, 313x129, 134x129) are asserted here as well.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import modules_ref as mr, stft_ref as sr, synth  # noqa: E402


def torch_stft(wav, n_fft, hop, window):
    w = torch.from_numpy(sr.get_window(window, n_fft))
    X = torch.stft(torch.from_numpy(wav), n_fft, hop, n_fft, w, center=True, pad_mode='reflect',
                   onesided=True, return_complex=True)
    return X.numpy()                                            # [B,F,T] complex128


def stft_case(name, L, hop, B, seed, window='hann'):
    b = synth.make_batch(B, L, 2, seed=seed)
    wav = b['mix_wav']
    X = torch_stft(wav, 256, hop, window)
    T = 1 + L // hop
    assert X.shape == (B, 129, T)
    # inverse with torch.istft on a masked spectrum (random sigmoid-like mask)
    rng = np.random.RandomState(seed)
    mask = rng.uniform(0, 1, X.shape)
    Y = torch.from_numpy(mask * X)
    w = torch.from_numpy(sr.get_window(window, 256))
    y = torch.istft(Y, 256, hop, 256, w, center=True, length=hop * (T - 1)).numpy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), wav=wav.astype(np.float64),
                        spec=np.transpose(X, (0, 2, 1)).astype(np.complex64),
                        mask=np.transpose(mask, (0, 2, 1)).astype(np.float32),
                        wav_out=y.astype(np.float32), hop=hop, window=window)


def model_case(name, cell, layers, cplx, B, T, S, seed):
    torch.manual_seed(seed)
    rc = mr.RefConfig(NUM_LAYERS=layers, is_ComlexMask=cplx)
    mix = mr.MIX_SPEECH(rc, 129, T, cell, layers)
    emb = mr.SPEECH_EMBEDDING(rc, 101, 50, 2)
    att = mr.ATTENTION(rc, 50, 'dot')
    adj = mr.ADDJUST(rc, 600, 50)
    feas = torch.rand(B, T, 129) * 3
    mag = torch.randn(B, T, 129, 2)
    idx = np.sort(np.random.RandomState(seed).choice(101, (B, S)), axis=1)
    y = torch.rand(B, S, T, 129, 2) if cplx else torch.rand(B, S, T, 129)
    with torch.no_grad():
        r = mr.forward_ref(rc, mix, emb, att, adj, feas, idx, mag)
        loss = mr.loss_ref(rc, r, y)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), feas=feas.numpy(), mag=mag.numpy(), idx=idx,
                        target=y.numpy(), masks=r['masks'].numpy(), hidden=r['hidden'].numpy(),
                        query=r['query'].numpy(), loss=np.array([float(v) for v in loss]),
                        cell=cell, layers=layers, cplx=cplx, seed=seed)


def main():
    # reference shape constants
    assert sr.num_frames(40000, 128) == 313 and 128 * (313 - 1) == 39936
    assert sr.num_frames(17040, 128) == 134
    stft_case('stft_hop128', 5120, 128, 2, 1)
    stft_case('stft_hop64', 3000, 64, 2, 2)
    stft_case('stft_sine', 4096, 128, 1, 3, 'sine')
    model_case('model_lstm2_real', 'lstm', 2, False, 2, 24, 2, 1)
    model_case('model_gru2_crm', 'gru', 2, True, 2, 20, 3, 2)
    model_case('model_lstm4_real', 'lstm', 4, False, 1, 16, 2, 3)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()

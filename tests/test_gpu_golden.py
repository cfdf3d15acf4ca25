"""GPU parity against the COMMITTED fixtures under tests/golden/ (read directly, no live oracle in between).

  ref_stft_*.npz   outputs of the reference's own numpy stft/istft (Cocktail/software/DL4SS_Keras/test_stft_istft.py:9-63,
                   written by tests/golden/make_ref_fixtures.py): K1 and K6 must reproduce them on every frame / sample
                   the centred (librosa) and un-centred (that file's) conventions share, within 1e-4 of the peak.
  stft_*.npz       torch.stft / torch.istft with the reference's librosa settings (make_golden.py).
  model_*.npz      the reference's torch arithmetic through oracle/modules_ref.py with seeded weights (make_golden.py):
                   masks / encoder output / speaker queries / losses.
"""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-4


def gold(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)


def peak_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize('name', ['ref_stft_256_2', 'ref_stft_256_4'])
def test_k1_k6_reproduce_reference_numpy_stft(cuda, name):
    """K1 / K6 against values computed by the reference's own code.  A centred frame t starts at t*hop - n_fft/2, so the
    reference's frame i (no padding, start i*hop) is centred frame i + overlap/2; its inverse overlap-adds frames
    0 .. n-overlap-1 untrimmed, K6 trims n_fft/2 at both ends of the same sum."""
    import dl4ss_b200 as d
    g = gold(name)
    n_fft, ov = int(g['n_fft']), int(g['overlap'])
    hop = n_fft // ov
    n = g['X'].shape[0]
    sine = [np.sin(i * np.pi / n_fft) for i in range(n_fft)]
    for window, key_x, key_y in ((g['window'], 'X', 'y'), (sine, 'Xs', 'ys')):
        for dtype in (torch.float64, torch.float32):
            wav = torch.from_numpy(g['x']).to(cuda, dtype).unsqueeze(0)
            feat, cplx = d.stft_features(wav, n_fft, hop, window, 'abs')
            got = torch.view_as_complex(cplx)[0].cpu().numpy()[ov // 2: ov // 2 + n]
            assert got.shape == g[key_x].shape
            assert peak_err(got, g[key_x]) < TOL
            assert peak_err(feat[0].cpu().numpy()[ov // 2: ov // 2 + n], np.abs(g[key_x])) < TOL
        # inverse of the reference's spectrum, per-source form (no mask)
        spec = torch.view_as_real(torch.from_numpy(g[key_x][:n - ov].astype(np.complex64))).to(cuda)
        y = d.mask_istft(None, spec.view(1, 1, n - ov, n_fft // 2 + 1, 2).contiguous(), hop, window)[0, 0].cpu().numpy()
        want = g[key_y][n_fft // 2: hop * (n - 1) - n_fft // 2]
        assert y.shape == want.shape
        assert peak_err(y, want) < TOL
    # masked reconstruction: real mask x mixture spectrum inside K6 == the reference's istft(mask * X)
    mask = torch.from_numpy(g['mask'][:n - ov]).to(cuda).view(1, 1, n - ov, -1).contiguous()
    spec = torch.view_as_real(torch.from_numpy(g['X'][:n - ov].astype(np.complex64))).to(cuda).unsqueeze(0).contiguous()
    ym = d.mask_istft(mask, spec, hop, g['window'])[0, 0].cpu().numpy()
    assert peak_err(ym, g['ym'][n_fft // 2: hop * (n - 1) - n_fft // 2]) < TOL
    # round trip through both kernels returns the reference's input waveform
    wav = torch.from_numpy(g['x']).to(cuda).unsqueeze(0)
    _, c = d.stft_features(wav, n_fft, hop, g['window'], None)
    T = c.shape[1]
    back = d.mask_istft(None, c.view(1, 1, T, -1, 2), hop, g['window'])[0, 0].cpu().numpy()
    m = back.shape[0]
    assert np.abs(back[hop:m - hop] - g['x'][hop:m - hop]).max() < TOL


@pytest.mark.parametrize('name', ['stft_hop128', 'stft_hop64', 'stft_sine'])
def test_k1_k6_match_torch_stft_fixtures(cuda, name):
    import dl4ss_b200 as d
    g = gold(name)
    hop, window = int(g['hop']), str(g['window'])
    wav = torch.from_numpy(g['wav']).to(cuda)
    feat, cplx = d.stft_features(wav, 256, hop, window, 'abs')
    assert peak_err(torch.view_as_complex(cplx).cpu().numpy(), g['spec']) < TOL
    assert peak_err(feat.cpu().numpy(), np.abs(g['spec'])) < TOL
    spec = torch.view_as_real(torch.from_numpy(g['spec'])).to(cuda).contiguous()
    mask = torch.from_numpy(g['mask']).to(cuda).unsqueeze(1).contiguous()
    y = d.mask_istft(mask, spec, hop, window)[:, 0].cpu().numpy()
    assert y.shape == g['wav_out'].shape
    assert peak_err(y, g['wav_out']) < TOL


@pytest.mark.parametrize('name', ['model_lstm2_real', 'model_gru2_crm', 'model_lstm4_real'])
def test_masks_match_model_fixtures(cuda, name):
    """Weights are re-created from the fixture's seed exactly as make_golden.py did (same constructor order under
    torch.manual_seed), inputs and expected outputs come from the file."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    g = gold(name)
    cell, layers, cplx, seed = str(g['cell']), int(g['layers']), bool(g['cplx']), int(g['seed'])
    B, T = g['feas'].shape[:2]
    torch.manual_seed(seed)
    rc = mr.RefConfig(NUM_LAYERS=layers, is_ComlexMask=cplx)
    ref = {'mix': mr.MIX_SPEECH(rc, 129, T, cell, layers), 'emb': mr.SPEECH_EMBEDDING(rc, 101, 50, 2),
           'att': mr.ATTENTION(rc, 50, 'dot'), 'adj': mr.ADDJUST(rc, 600, 50)}
    old = (d.config.NUM_LAYERS, d.config.is_ComlexMask, d.config.is_SelfTune)
    d.config.HIDDEN_UNITS, d.config.EMBEDDING_SIZE, d.config.NUM_LAYERS = 300, 50, layers
    d.config.is_ComlexMask, d.config.is_SelfTune = cplx, True
    try:
        ours = {'mix': d.MIX_SPEECH(129, T, cell=cell, num_layers=layers).to(cuda), 'emb': d.SPEECH_EMBEDDING(101, 50, 2).to(cuda),
                'att': d.ATTENTION(50, 'dot').to(cuda), 'adj': d.ADDJUST(600, 50).to(cuda)}
        for k in ours:
            ours[k].load_state_dict(copy.deepcopy(ref[k].state_dict()))
        sep = d.Separator(ours['mix'], ours['emb'], ours['att'], ours['adj'])
        feas = torch.from_numpy(g['feas']).to(cuda)
        idx = g['idx']
        extras = {}
        hidden = ours['mix'].encode(feas, extras)
        assert np.abs(hidden.cpu().numpy() - g['hidden']).max() < 2e-5
        q, _ = sep.queries(hidden, idx, extras.get('hmean'))
        assert np.abs(q.cpu().numpy() - g['query']).max() < 2e-5
        m = sep.masks(feas, idx)
        want = g['masks']
        assert tuple(m.shape) == want.shape
        if not cplx:
            assert np.abs(m.cpu().numpy() - want).max() < TOL
            mix, tgt = feas, torch.from_numpy(g['target']).to(cuda)
        else:
            err = np.abs(m.cpu().numpy() - want) / np.maximum(np.abs(want), 1.0 / d.config.cRM_C)
            assert err[np.abs(want) < 60.0].max() < TOL          # the decompression's conditioning
            mix, tgt = torch.from_numpy(g['mag']).to(cuda), torch.from_numpy(g['target']).to(cuda)
        loss = d.mask_loss(m, mix, tgt)
        for got, exp in zip(loss, g['loss']):
            assert abs(got.item() - exp) < 1e-4 * abs(exp) + 1e-7
    finally:
        d.config.NUM_LAYERS, d.config.is_ComlexMask, d.config.is_SelfTune = old

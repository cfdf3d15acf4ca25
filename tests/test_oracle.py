"""CPU tests that pin the oracle (run with -m "not gpu").

The reference holds no value fixtures and its STFT/SDR libraries are absent, so the oracle
is pinned against (i) the committed golden vectors under tests/golden/ (made by make_golden.py from
torch.stft/istft and from torch.nn -- the reference's own arithmetic), (ii) the shape constants the
reference embeds, (iii) analytic identities.
"""
import os

import numpy as np
import pytest
import torch

from oracle import bss_eval_ref as be
from oracle import modules_ref as mr
from oracle import stft_ref as sr
from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)


# ------------------------------------------------------------------------------ STFT / iSTFT
def test_reference_shape_constants():
    # TDAA_beta/main_run_sstune_EvalVer.py:49 (39936), Discriminator Linear(36480,1) -> T=313, F=129
    assert sr.stft_ref(np.zeros(40000) + 1e-3, 256, 128).shape == (129, 313)
    assert sr.istft_ref(np.zeros((129, 313), np.complex64), 128).shape == (39936,)
    # Torch_multi/predata_multiAims.py:58,67: (5,17040) -> (5,134,129)
    assert sr.stft_ref(np.ones(17040), 256, 128).T.shape == (134, 129)
    # hop 64 (BASELINE configs[0]): 4 s -> 501 frames
    assert sr.num_frames(32000, 64) == 501


@pytest.mark.parametrize('name', ['stft_hop128', 'stft_hop64', 'stft_sine'])
def test_stft_oracle_matches_golden(name):
    g = gold(name)
    hop, window = int(g['hop']), str(g['window'])
    for b in range(g['wav'].shape[0]):
        S = sr.stft_ref(g['wav'][b], 256, hop, window).T
        assert S.dtype == np.complex64
        assert np.abs(S - g['spec'][b]).max() < 1e-5 * np.abs(g['spec'][b]).max()
        y = sr.istft_ref((g['mask'][b] * g['spec'][b]).T, hop, window)
        assert y.dtype == np.float32 and y.shape == g['wav_out'][b].shape
        assert np.abs(y - g['wav_out'][b]).max() < 1e-5 * np.abs(g['wav_out'][b]).max()


@pytest.mark.parametrize('hop,L', [(128, 4000), (64, 2000), (100, 1234), (256, 2048)])
def test_stft_oracle_matches_torch_live(hop, L):
    rng = np.random.RandomState(L)
    y = rng.standard_normal(L)
    w = torch.from_numpy(sr.get_window('hann', 256))
    X = torch.stft(torch.from_numpy(y), 256, hop, 256, w, center=True, pad_mode='reflect', return_complex=True).numpy()
    S = sr.stft_ref(y, 256, hop)
    assert S.shape == X.shape == (129, 1 + L // hop)
    assert np.abs(S - X).max() < 1e-5 * np.abs(X).max()
    if 256 % hop == 0 and hop < 256:   # window sum-square > 0 everywhere: the round trip is the identity
        back = sr.istft_ref(S, hop)
        assert np.abs(back - y[:back.shape[0]]).max() < 1e-5 * np.abs(y).max()


def test_stft_conj_and_windows():
    y = np.random.RandomState(0).standard_normal(3000)
    assert np.array_equal(sr.stft_ref(y, conj=True), np.conj(sr.stft_ref(y)))
    w = sr.get_window('hann', 256)
    from scipy.signal import get_window
    assert np.allclose(w, get_window('hann', 256, fftbins=True), atol=1e-15)
    assert np.allclose(sr.get_window('sine', 256), [np.sin(i * np.pi / 256) for i in range(256)])
    with pytest.raises(ValueError):
        sr.stft_ref(np.zeros(100))             # reflect pad needs L > n_fft/2
    with pytest.raises(ValueError):
        sr.get_window(np.ones(5), 256)


def test_features_ref_layouts():
    y = np.random.RandomState(3).standard_normal(5000)
    f = sr.features_ref(y, 256, 128)
    T = 1 + 5000 // 128
    assert f['mix_feas'].shape == (T, 129) and f['mix_feas'].dtype == np.float32
    assert f['mix_phase'].shape == (T, 129) and f['mix_phase'].dtype == np.complex64
    assert f['mix_mag'].shape == (T, 129, 2)
    assert np.array_equal(f['mix_mag'][..., 0], f['mix_phase'].real)       # convert2
    assert np.allclose(f['mix_feas'], np.abs(f['mix_phase']))
    fl = sr.features_ref(np.zeros(5000), 256, 128, log_spectral=True)
    assert np.allclose(fl['mix_feas'], np.log(np.spacing(1)))              # digital silence -> log(eps)


def test_preprocess_source():
    x = np.random.RandomState(1).standard_normal(1000) + 3.0
    p = sr.preprocess_source(x, 1500, 0.0)
    assert p.shape == (1500,) and abs(np.abs(p).max() - 1.0) < 1e-12 and np.all(p[1000:] == 0)
    assert abs(p[:1000].mean()) < 0.05
    p2 = sr.preprocess_source(x, 800, 6.0)
    assert p2.shape == (800,) and abs(np.abs(p2).max() - 10 ** 0.3) < 1e-12


# ------------------------------------------------------------------------------ model stages
@pytest.mark.parametrize('name', ['model_lstm2_real', 'model_gru2_crm', 'model_lstm4_real'])
def test_model_oracle_matches_golden(name):
    g = gold(name)
    cell, layers, cplx, seed = str(g['cell']), int(g['layers']), bool(g['cplx']), int(g['seed'])
    T = g['feas'].shape[1]
    torch.manual_seed(seed)
    rc = mr.RefConfig(NUM_LAYERS=layers, is_ComlexMask=cplx)
    mods = (mr.MIX_SPEECH(rc, 129, T, cell, layers), mr.SPEECH_EMBEDDING(rc, 101, 50, 2),
            mr.ATTENTION(rc, 50, 'dot'), mr.ADDJUST(rc, 600, 50))
    with torch.no_grad():
        r = mr.forward_ref(rc, *mods, torch.from_numpy(g['feas']), g['idx'], torch.from_numpy(g['mag']))
        loss = mr.loss_ref(rc, r, torch.from_numpy(g['target']))
    # same torch build -> same bits up to thread-count dependent summation order
    assert np.abs(r['hidden'].numpy() - g['hidden']).max() < 1e-5
    assert np.abs(r['query'].numpy() - g['query']).max() < 1e-5
    scale = np.maximum(np.abs(g['masks']), 10.0 if cplx else 1.0)
    ok = np.abs(g['masks']) < 60.0
    assert (np.abs(r['masks'].numpy() - g['masks']) / scale)[ok].max() < 1e-4
    if not cplx:
        assert abs(float(loss[0]) - g['loss'][0]) < 1e-5 * abs(g['loss'][0])


def test_attention_dot_is_inner_product():
    rc = mr.RefConfig()
    att = mr.ATTENTION(rc, 50, 'dot')
    emb = torch.randn(3, 7, 129, 50)
    q = torch.randn(3, 50)
    m = att(emb, q)
    assert torch.allclose(m, torch.sigmoid(torch.einsum('ntfe,ne->ntf', emb, q)), atol=1e-6)
    rc2 = mr.RefConfig(is_ComlexMask=True)
    att2 = mr.ATTENTION(rc2, 50, 'dot')
    q2 = torch.randn(3, 100)
    m2 = att2(emb, q2)
    assert m2.shape == (3, 7, 129, 2)
    assert torch.allclose(m2[..., 1], 10 * torch.tanh(torch.einsum('ntfe,ne->ntf', emb, q2[:, 50:])), atol=1e-5)
    with pytest.raises(IndexError):
        mr.ATTENTION(rc2, 50, 'align')(emb, q2)
    with pytest.raises(IndexError):
        mr.ATTENTION(rc, 50, 'nope')(emb, q)


def test_crm_decompression_closed_form():
    # -1/C*log((K-K*tanh e)/(K+K*tanh e)) == 2e/C in exact arithmetic (cRM numerics)
    e = torch.linspace(-3, 3, 61, dtype=torch.float64)
    m = mr.cRM_k * torch.tanh(e)
    M = -1 / mr.cRM_C * torch.log((mr.cRM_k - m) / (mr.cRM_k + m))
    assert torch.allclose(M, 2 * e / mr.cRM_C, atol=1e-9)


def test_top_k_mask_and_multihot_embedding():
    p = torch.tensor([[0.1, 0.9, 0.8, 0.2], [0.7, 0.1, 0.2, 0.95]])
    m = mr.top_k_mask(p, 0.5, 2)
    assert m.tolist() == [[0, 1, 1, 0], [1, 0, 0, 1]]
    m = mr.top_k_mask(p, 0.85, 2)
    assert m.tolist() == [[0, 1, 0, 0], [0, 0, 0, 1]]
    torch.manual_seed(0)
    old = mr.SPEECH_EMBEDDING_multihot(4, 6, 2)
    new = mr.SPEECH_EMBEDDING(mr.RefConfig(EMBEDDING_SIZE=6), 4, 6, 2)
    new.layer.weight.data.copy_(old.layer.weight.data)
    full = old(mr.top_k_mask(p, 0.5, 2))
    act = new(None, [[1, 2], [0, 3]])
    assert torch.equal(full[0, [1, 2]], act[0]) and torch.equal(full[1, [0, 3]], act[1])
    assert torch.all(full[0, [0, 3]] == 0)


def test_pit_oracle_recovers_permutation():
    t = torch.randn(4, 3, 10, 5)
    perm = [2, 0, 1]
    p = t[:, perm] + 0.01 * torch.randn(4, 3, 10, 5)
    loss, best = mr.pit_mse_ref(p, t)
    assert loss.item() < 1e-3
    assert all(best[b].tolist() == perm for b in range(4))


# ------------------------------------------------------------------------------ SDR
def test_bss_eval_identities():
    rng = np.random.RandomState(0)
    n = 6000
    s = np.stack([synth.speech_like(rng, n), synth.speech_like(rng, n)])
    sdr, sir, sar, perm = be.bss_eval_sources(s, s[::-1] * 0.5)          # scaled + swapped copies
    assert list(perm) == [1, 0] and sdr.min() > 60.0
    mix = s.sum(0)
    sdr, _, _, _ = be.bss_eval_sources(s, np.stack([mix, mix]))
    p = (s ** 2).sum(1)
    expect = 10 * np.log10(p / p[::-1])                                   # SDR of the mixture vs each source
    assert np.abs(sdr - expect).max() < 1.5
    noisy = s + 0.1 * rng.standard_normal(s.shape) * s.std()
    sdr, _, _, perm = be.bss_eval_sources(s, noisy)
    assert list(perm) == [0, 1] and 15.0 < sdr.mean() < 25.0


def test_synth_batch_is_reference_shaped():
    b = synth.make_batch(3, 8000, 2, seed=5, active_len=(3000, 8000))
    assert b['mix_wav'].shape == (3, 8000) and b['sources'].shape == (3, 2, 8000)
    assert np.allclose(b['mix_wav'], b['sources'].sum(1))
    assert np.all(np.diff(b['spk_idx'], axis=1) > 0)                      # sorted speaker order
    peak = np.abs(b['sources']).max(-1)
    assert np.allclose(peak, 10 ** (b['gains_db'] / 20.0))
    b2 = synth.make_batch(3, 8000, 2, seed=5, active_len=(3000, 8000))
    assert np.array_equal(b['mix_wav'], b2['mix_wav'])


def test_discriminator_oracle_shape_constant_and_losses():
    """The reference's own shape constant: Linear(36480, 1) after three 3x3/stride-2 convolutions of a [313,129]
    spectrogram (TDAA_beta/main_run_sstune_EvalVer.py:334), and the least-squares GAN terms of :643-652,670."""
    import dl4ss_b200.modules as M
    assert M.Discriminator.flat_features(313, 129) == 36480
    torch.manual_seed(3)
    dis = mr.Discriminator()
    with torch.no_grad():
        s_true = dis(torch.rand(2, 2, 313, 129))
        s_false = dis(torch.rand(2, 2, 313, 129) * 0.5)
    assert s_true.shape == (4, 1) and float(s_true.min()) > 0 and float(s_true.max()) < 1
    ref = mr.gan_loss_terms_ref(s_true, s_false)
    ours = M.gan_loss_terms(s_true, s_false)          # pure tensor arithmetic: runs on the CPU too
    for k in ('loss_dis_true', 'loss_dis_false', 'loss_dis', 'loss_gen', 'acc_true', 'acc_false', 'acc_dis'):
        assert abs(float(ours[k]) - float(ref[k])) < 1e-6, k
    assert abs(float(ref['loss_dis']) - float(((s_true - 1) ** 2).mean() + (s_false ** 2).mean())) < 1e-7

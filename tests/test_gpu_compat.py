"""GPU tests of the reference-named loop-level entry points (dl4ss_b200/compat.py) against the oracle:
prepare_data's yield protocol and values, bss_eval's wav files + bss_test.cal's SDR, the eval_bss epoch end to end,
the old multi-hot SPEECH_EMBEDDING / `multi_mask * top_k_mask` form, the AUGMENT_DATA circular shift."""
import copy
import os
import random

import numpy as np
import pytest
import torch

from tests.util import rel_err, build_pair

pytestmark = pytest.mark.gpu


@pytest.fixture
def small_config(monkeypatch):
    import dl4ss_b200 as d
    for k, v in dict(BATCH_SIZE=4, MAX_LEN=16000, MIN_MIX=2, MAX_MIX=2, AUGMENT_DATA=False, SHUFFLE_BATCH=False,
                     is_ComlexMask=False, is_SelfTune=True, Out_Sep_Result=True, IS_LOG_SPECTRAL=False).items():
        monkeypatch.setattr(d.config, k, v)
    monkeypatch.setattr(d.compat, 'log', lambda *a: None)
    d.compat.set_source(d.compat.SyntheticSource(num_mixtures=9, seed=5))
    yield d
    d.compat.set_source(None)


def _oracle_mixture(src, split, recipe, L, shifts=None):
    """The reference's per-source preprocessing in float64 (oracle/stft_ref.py preprocess_source) + the optional shift."""
    from oracle import stft_ref as sr
    spk, db, names = recipe
    outs = []
    for k in range(len(spk)):
        x, _ = src.read(split, spk[k], names[k])
        x = np.array(x[:L], dtype=np.float64)
        if shifts is not None:                       # predata_fromList.py:150-153 sits between the normalisation and the padding
            x = x - x.mean()
            x = x / np.abs(x).max()
            x = np.append(x[shifts[k]:], x[:shifts[k]])
            x = np.append(x, np.zeros(L - len(x))) * 10.0 ** (db[k] / 20.0)
        else:
            x = sr.preprocess_source(x, L, db[k])
        outs.append(x)
    return np.stack(outs)


@pytest.mark.parametrize('cplx', [False, True])
def test_prepare_data_protocol_and_values(cuda, small_config, cplx):
    d = small_config
    from oracle import stft_ref as sr
    d.config.is_ComlexMask = cplx
    src = d.compat.get_source()
    g = next(d.prepare_data('global', 'train'))
    all_spk, s2i, i2s, T, F, frames, num_spk, batch_total = g
    assert (T, F, frames, num_spk, batch_total) == (126, 129, 32, 101, 2) and all_spk == sorted(all_spk)
    assert all(i2s[s2i[s]] == s for s in all_spk)
    random.seed(1)
    gen = d.prepare_data('once', 'valid')
    recipes = src.recipes('valid', 2)
    n_batches = 0
    while True:
        data = next(gen)
        if data is False:
            break
        keys = {'mix_wav', 'mix_feas', 'mix_phase', 'aim_fea', 'aim_spkname', 'query', 'num_all_spk',
                'multi_spk_fea_list', 'multi_spk_wav_list', 'batch_total'} | ({'mix_mag'} if cplx else set())
        assert set(data.keys()) == keys
        assert data['mix_wav'].shape == (4, 16000) and data['mix_wav'].dtype == np.float64
        assert data['mix_feas'].shape == (4, 126, 129) and data['mix_feas'].dtype == np.float32
        assert data['mix_phase'].dtype == np.complex64 and data['num_all_spk'] == 101 and data['batch_total'] == 2
        for b in range(4):
            rec = recipes[n_batches * 4 + b]
            want = _oracle_mixture(src, 'valid', rec, 16000)
            assert list(data['multi_spk_wav_list'][b].keys()) == rec[0] and data['aim_spkname'][b] == rec[0][0]
            for k, spk in enumerate(rec[0]):
                assert np.abs(data['multi_spk_wav_list'][b][spk] - want[k]).max() < 4e-6 * np.abs(want[k]).max()
                S = sr.stft_ref(want[k], 256, 128).T
                tgt = sr.convert2(S) if cplx else np.abs(S)
                assert rel_err(data['multi_spk_fea_list'][b][spk], tgt) < 1e-4
            f = sr.features_ref(want.sum(0), 256, 128)
            assert np.abs(data['mix_wav'][b] - want.sum(0)).max() < 8e-6 * np.abs(want).max()
            assert rel_err(data['mix_feas'][b], f['mix_feas']) < 1e-4
            assert rel_err(data['mix_phase'][b], f['mix_phase']) < 1e-4
            if cplx:
                assert rel_err(data['mix_mag'][b], f['mix_mag']) < 1e-4
            assert np.array_equal(data['aim_fea'][b], data['multi_spk_fea_list'][b][rec[0][0]])
        n_batches += 1
    assert n_batches == 2                                    # 9 mixtures // BATCH_SIZE 4, then False


def test_prepare_data_augment_shift(cuda, small_config):
    """AUGMENT_DATA on the training split: the circular shift drawn with random.sample(range(len(signal)), 1)."""
    d = small_config
    d.config.AUGMENT_DATA = True
    src = d.compat.get_source()
    random.seed(7)
    data = next(d.prepare_data('once', 'train'))
    random.seed(7)                                           # replay the generator's draws: mix_k, then one shift per source
    random.randint(2, 2)
    recipes = src.recipes('train', 2)
    for b in range(4):
        shifts = []
        for k in range(2):
            n = min(len(src.read('train', recipes[b][0][k], recipes[b][2][k])[0]), 16000)
            shifts.append(random.sample(range(n), 1)[0])
        want = _oracle_mixture(src, 'train', recipes[b], 16000, shifts)
        for k, spk in enumerate(recipes[b][0]):
            assert np.abs(data['multi_spk_wav_list'][b][spk] - want[k]).max() < 4e-6 * np.abs(want[k]).max()
    assert max(shifts) > 0


def _decisive_classifier(d, mr, cfg, T, seed):
    torch.manual_seed(seed)
    cls_ref = mr.MIX_SPEECH_classifier(cfg, 129, T, 101)
    with torch.no_grad():
        cls_ref.Linear.weight.mul_(40.0)
        cls_ref.Linear.bias[0] = -50.0                      # speaker index 0 reads as 'channel off' in bss_eval (EvalVer.py:57)
    cls = d.MIX_SPEECH_classifier(129, T, 101).cuda()
    cls.load_state_dict(copy.deepcopy(cls_ref.state_dict()))
    return cls_ref, cls


def test_eval_bss_epoch_matches_oracle(cuda, small_config, tmp_path, monkeypatch):
    """The reference's evaluation epoch (EvalVer.py:407-507) through compat.eval_bss: classifier -> top_k_mask -> embedding
    + ADDJUST -> attention masks -> bss_eval wav files -> bss_test.cal, against the same flow on the oracle (torch-CPU
    modules, numpy istft, PCM16 round trip, mir_eval restatement): SDR of every source within 0.01 dB."""
    d = small_config
    from oracle import modules_ref as mr, stft_ref as sr, bss_eval_ref as be
    monkeypatch.chdir(tmp_path)
    T = 126
    ref, ours = build_pair('lstm', 2, 129, T, False)
    cls_ref, cls = _decisive_classifier(d, mr, ref['cfg'], T, 13)
    all_spk, s2i, i2s = next(d.prepare_data('global', 'train'))[:3]
    sdr = d.eval_bss(ours['mix'], ours['adj'], cls, ours['emb'], ours['att'], torch.nn.MSELoss(), s2i, i2s, 101, T, 129)
    assert sdr.shape == (16,)
    files = sorted(os.listdir('batch_output'))
    assert len([f for f in files if f.endswith('_pre.wav')]) == 8 and '0_True_mix.wav' in files     # the last batch's files
    # the same epoch on the oracle
    want = []
    gen = d.prepare_data('once', 'valid')
    while True:
        data = next(gen)
        if data is False:
            break
        feas = torch.from_numpy(data['mix_feas'])
        with torch.no_grad():
            sel = mr.top_k_mask(cls_ref(feas), -0.5, 2)
            idx = np.stack([np.where(line == 1)[0] for line in sel.numpy()])
            assert (idx > 0).all()
            r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, idx)
        wav = mr.reconstruct_ref(r, data['mix_phase'], 128)
        for b in range(4):
            names = sorted(data['multi_spk_wav_list'][b].keys())
            real = np.stack([data['multi_spk_wav_list'][b][s][:39936] for s in names])
            order = np.argsort([i2s[int(i)] for i in idx[b]])
            pre = wav[b][order]
            n = min(real.shape[1], pre.shape[1])
            want.append(be.bss_eval_sources(be.pcm16_roundtrip(real[:, :n]), be.pcm16_roundtrip(pre[:, :n]))[0])
    want = np.concatenate(want)
    assert np.abs(sdr - want).max() < 0.01, (sdr, want)
    # in-memory scoring of the last batch == scoring its wav files
    a = d.bss_test.cal('batch_output/', 2)
    b = d.bss_test.cal(None, 2)
    assert np.abs(a - b).max() < 1e-6 and np.abs(a - sdr[-8:]).max() < 1e-6
    f = d.bss_test.cal(None, 2, pcm16=False)
    assert np.abs(f - a).max() < 0.05                       # the float number is close to, not equal to, the PCM16 one


def test_bss_eval_crm_files(cuda, small_config, tmp_path, monkeypatch):
    d = small_config
    from oracle import stft_ref as sr
    monkeypatch.chdir(tmp_path)
    d.config.is_ComlexMask = True
    data = next(d.prepare_data('once', 'valid'))
    i2s = next(d.prepare_data('global', 'train'))[2]
    s2i = {v: k for k, v in i2s.items()}
    B, T = 4, 126
    rng = np.random.RandomState(0)
    pr = rng.standard_normal((B, 2, T, 129)).astype(np.float32)
    pi = rng.standard_normal((B, 2, T, 129)).astype(np.float32)
    y = np.stack([[data['multi_spk_fea_list'][b][s] for s in data['multi_spk_fea_list'][b]] for b in range(B)])
    idx = [[s2i[s] for s in data['multi_spk_fea_list'][b]] for b in range(B)]
    out = d.bss_eval_cRM(torch.from_numpy(pr).cuda(), torch.from_numpy(pi).cuda(), torch.from_numpy(y).cuda(), idx, i2s, data)
    for b in range(B):
        for k, spk in enumerate(data['multi_spk_fea_list'][b]):
            if idx[b][k] == 0:
                continue
            want = sr.istft_ref((pr[b, k] + 1j * pi[b, k]).T, 128)
            got, rate = d.compat.read_wav_pcm('batch_output/{}_{}_pre.wav'.format(b, spk))
            assert rate == 8000 and np.abs(got - np.round(np.clip(want, -1, 1 - 2.0 ** -15) * 32768) / 32768).max() <= 1.01 / 32768
            gen, _ = d.compat.read_wav_pcm('batch_output/{}_{}_genTrue.wav'.format(b, spk))
            src = np.clip(data['multi_spk_wav_list'][b][spk][:len(gen)], -1.0, 1.0 - 2.0 ** -15)   # PCM16 clips the +2.5 dB peaks
            assert np.abs(gen[128:-128] - src[128:-128]).max() < 2.0 / 32768          # iSTFT(STFT(source)) == source
    assert len(out.samples) == B


def test_multihot_embedding_and_masks(cuda):
    """Torch_multi/main_run_multi_selfSS.py:308-328,476-493: SPEECH_EMBEDDING(top_k_mask) and multi_mask * top_k_mask."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, T, N = 3, 21, 10
    ref, ours = build_pair('lstm', 2, 129, T, False, self_tune=False, num_spk=N)
    try:
        torch.manual_seed(3)
        emb_ref = mr.SPEECH_EMBEDDING_multihot(N, 50, 2)
        emb_ref.load_state_dict(copy.deepcopy(ref['emb'].state_dict()))
        sel = torch.zeros(B, N)
        sel[0, [1, 4]] = 1
        sel[1, [0, 9]] = 1
        sel[2, [2, 3, 5, 6, 7]] = 1                         # more than one fused-kernel group of 4
        with torch.no_grad():
            e_ref = emb_ref(sel)
            e = ours['emb'](sel.cuda())
        assert tuple(e.shape) == (B, N, 50) and (e.cpu() - e_ref).abs().max().item() == 0.0
        feas = torch.rand(B, T, 129) * 2
        with torch.no_grad():
            hid = ref['mix'](feas)[0]
            h5 = hid.view(B, 1, T, 129, 50).expand(B, N, T, 129, 50).contiguous().view(-1, T, 129, 50)
            want = ref['att'](h5, e_ref.view(-1, 50)).view(B, N, T, 129) * sel.view(B, N, 1, 1)
        sep = d.Separator(ours['mix'], ours['emb'], ours['att'], None)
        got = sep.masks_multihot(feas.cuda(), sel.cuda()).cpu()
        assert (got - want).abs().max().item() < 1e-4
        assert (got[sel == 0] == 0).all()
        # module-level form (the reference glue) gives the same thing
        with torch.no_grad():
            o, _ = ours['mix'](feas.cuda())
            h5 = o.view(B, 1, T, 129, 50).expand(B, N, T, 129, 50).contiguous().view(-1, T, 129, 50)
            got2 = ours['att'](h5, e.view(-1, 50)).view(B, N, T, 129) * sel.cuda().view(B, N, 1, 1)
        assert (got2.cpu() - want).abs().max().item() < 1e-4
    finally:
        d.config.is_SelfTune = True

"""Oracle-compared parity at the BASELINE configurations' full depth ("What's weak" 3-5).

  configs[1]  TDAA_beta LSTM 4x300 + ADDJUST, dot attention, 5 s utterances (T = 313), S = 2: B = 40 crosses one
              32-utterance tile of the recurrent kernel and leaves a partial one.  Masks, spectra, waveforms <= 1e-4,
              SDR within 0.01 dB on 8 utterances, also after the reference's PCM16 wav round trip
              (TDAA_beta/main_run_sstune_EvalVer.py:420-470,55-72).
  configs[2]  GRU 2x300, cRM complex masks (2E = 100 wide queries), S = 3, T = 313, B = 260: crosses the 256-utterance
              launch boundary of the recurrent kernel; utterances are independent, so the oracle runs on the rows at
              both sides of every boundary {0, 31, 32, 255, 256, 259} (TDAA_beta/main_run_sstune_cRM_EvalVer.py:498-553).
  n1          MIX_SPEECH_classifier (BLSTM 3x600) at T = 313, B = 17 (one full 16-utterance tile + 1).
  cRM overflow  |energy| beyond tanh saturation: -1/C*log((K-m)/(K+m)) is +-inf in the reference
              (TDAA_beta/main_run_sstune_cRM_EvalVer.py:512) and must be +-inf here, fused and un-fused.
"""
import copy

import numpy as np
import pytest
import torch

from tests.util import rel_err, build_pair

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _oracle_features(wavs, hop=128):
    from oracle import stft_ref as sr
    feats = [sr.features_ref(w, 256, hop) for w in wavs]
    return (torch.from_numpy(np.stack([f['mix_feas'] for f in feats])),
            np.stack([f['mix_phase'] for f in feats]),
            torch.from_numpy(np.stack([f['mix_mag'] for f in feats])))


def test_config1_lstm4_full_depth_masks_waveforms_sdr(cuda):
    import dl4ss_b200 as d
    from oracle import modules_ref as mr, synth, bss_eval_ref as be
    B, L, S, hop = 40, 40000, 2, 128
    T = 1 + L // hop
    assert T == 313
    batch = synth.make_batch(B, L, S, seed=21)
    ref, ours = build_pair('lstm', 4, 129, T, False)
    feas, phase, _ = _oracle_features(batch['mix_wav'])
    with torch.no_grad():
        r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, batch['spk_idx'])
    wav_ref = mr.reconstruct_ref(r, phase, hop)
    sep = d.Separator(ours['mix'], ours['emb'], ours['att'], ours['adj'])
    out = sep.separate(torch.from_numpy(batch['mix_wav']).to(cuda), batch['spk_idx'], return_all=True)
    assert rel_err(out['mix_feas'].cpu().numpy(), feas.numpy()) < TOL
    assert rel_err(torch.view_as_complex(out['mix_mag']).cpu().numpy(), phase) < TOL
    m, rm = out['masks'].cpu(), r['masks']
    e_mask = (m - rm).abs().max().item()
    assert e_mask < TOL, e_mask
    big = rm > 1e-2
    assert ((m - rm).abs() / rm)[big].max().item() < TOL
    wav = out['wav'].cpu().numpy()
    assert wav.shape == wav_ref.shape == (B, S, 39936)
    e_wav = rel_err(wav, wav_ref)
    assert e_wav < TOL, e_wav
    n = wav.shape[-1]
    worst = worst_q = 0.0
    for b in (0, 1, 15, 16, 31, 32, 38, 39):                       # both sides of the tile boundary + the partial tile
        truth = batch['sources'][b][:, :n]
        s_ref = be.bss_eval_sources(truth, wav_ref[b])[0]
        s_got = be.bss_eval_sources(truth, wav[b])[0]
        worst = max(worst, float(np.abs(s_ref - s_got).max()))
        # "as-reference" number: every waveform goes through sf.write/sf.read PCM16 before scoring (EvalVer.py:72, bss_test.py:34)
        q = be.pcm16_roundtrip
        sq_ref = be.bss_eval_sources(q(truth), q(wav_ref[b]))[0]
        sq_got = be.bss_eval_sources(q(truth), q(wav[b]))[0]
        worst_q = max(worst_q, float(np.abs(sq_ref - sq_got).max()))
    print('configs[1] full depth: mask err %.2e, wav err %.2e, |dSDR| %.2e dB, PCM16 |dSDR| %.2e dB' % (e_mask, e_wav, worst, worst_q))
    assert worst < 0.01 and worst_q < 0.01
    # the on-device scorer agrees with the oracle's on the GPU output (float and PCM16 forms)
    from dl4ss_b200 import metrics
    rows = [0, 16, 39]
    truth = torch.from_numpy(batch['sources'][rows][:, :, :n].astype(np.float32)).to(cuda)
    est = out['wav'][rows]
    for f_gpu, f_cpu in ((lambda x: x, lambda x: x), (metrics.pcm16_roundtrip, be.pcm16_roundtrip)):
        sdr = metrics.bss_eval_sources_batch(f_gpu(truth), f_gpu(est))[0].cpu().numpy()
        for k in range(len(rows)):
            t = f_cpu(truth[k].cpu().numpy().astype(np.float64))
            e = f_cpu(est[k].cpu().numpy().astype(np.float64))
            assert np.abs(sdr[k] - be.bss_eval_sources(t, e)[0]).max() < 0.01


def test_config2_gru_crm_three_speakers_across_launch_boundary(cuda):
    import dl4ss_b200 as d
    from oracle import modules_ref as mr, synth
    B, L, S, hop = 260, 40000, 3, 128
    T = 1 + L // hop
    rows = [0, 31, 32, 255, 256, 259]
    batch = synth.make_batch(B, L, S, seed=22)
    ref, ours = build_pair('gru', 2, 129, T, True)
    try:
        # nn.Embedding's N(0,1) init with 2E = 100 wide queries puts ~5 % of the energies beyond |e| = 3, where the
        # decompression amplifies the last ulp of tanh (the saturated end is test_crm_overflow's subject):
        # scale the speaker table so that the masks sit in the range a trained cRM model uses
        with torch.no_grad():
            ref['emb'].layer.weight.mul_(0.25)
            ours['emb'].layer.weight.mul_(0.25)
        feas, _, mag = _oracle_features(batch['mix_wav'][rows])
        with torch.no_grad():
            r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, batch['spk_idx'][rows], mag)
        wav_ref = mr.reconstruct_ref(r, None, hop, complex_mask=True)
        sep = d.Separator(ours['mix'], ours['emb'], ours['att'], ours['adj'])
        out = sep.separate(torch.from_numpy(batch['mix_wav']).to(cuda), batch['spk_idx'], return_all=True)
        assert tuple(out['masks'].shape) == (B, S, T, 129, 2)
        assert rel_err(out['mix_mag'][rows].cpu().numpy(), mag.numpy()) < TOL
        m, rm = out['masks'][rows].cpu(), r['masks']
        err = (m - rm).abs() / torch.clamp(rm.abs(), min=1.0 / d.config.cRM_C)
        well = rm.abs() < 60.0
        assert well.float().mean().item() > 0.99                     # the well-conditioned range is (almost) everything
        e_mask = err[well].max().item()
        assert e_mask < TOL, e_mask
        assert err[rm.abs() < 140.0].max().item() < 1e-2
        e_wav = rel_err(out['wav'][rows].cpu().numpy(), wav_ref)
        print('configs[2] B=260: cRM mask err %.2e (|M|<60: %.4f of the bins), wav err %.2e, max|M| %.1f'
              % (e_mask, well.float().mean().item(), e_wav, rm.abs().max().item()))
        assert e_wav < TOL, e_wav
        assert not torch.isnan(out['wav']).any().item()
    finally:
        d.config.is_ComlexMask = 0


def test_classifier_full_length(cuda):
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, T = 17, 313
    ref, _ = build_pair('lstm', 1, 129, T, False)
    torch.manual_seed(31)
    cls_ref = mr.MIX_SPEECH_classifier(ref['cfg'], 129, T, 101)
    cls = d.MIX_SPEECH_classifier(129, T, 101).to(cuda)
    cls.load_state_dict(copy.deepcopy(cls_ref.state_dict()))
    feas = torch.rand(B, T, 129) * 2
    with torch.no_grad():
        p_ref = cls_ref(feas)
        y_ref, _ = cls_ref.layer(feas)
        p = cls(feas.to(cuda))
        y = d.rnn_forward(cls._packed, feas.to(cuda))
    e_y = (y.cpu() - y_ref).abs().max().item()
    e_p = (p.cpu() - p_ref).abs().max().item()
    print('classifier 3x600, T=313, B=17: BLSTM output err %.2e, probability err %.2e' % (e_y, e_p))
    assert e_y < 5e-5 and e_p < 2e-5
    assert torch.equal(d.top_k_mask(p, -0.5, 2).cpu(), mr.top_k_mask(p_ref, -0.5, 2))


def test_crm_overflow_region_is_inf_like_the_reference(cuda):
    """Energies far beyond tanh saturation: K*tanh(e) == +-K exactly in fp32, so the decompression's log argument is 0 or
    +inf and M = +-inf in the reference (cRM_EvalVer.py:512).  Both our forms must give the same infinities, and finite
    values for moderate energies.  The band 8 < |e| < 10.5, where the LAST ulp of tanh decides between a huge finite
    value and inf, is left out (either answer is 'the reference' there, depending on its libm)."""
    import dl4ss_b200 as d
    from dl4ss_b200 import _lib
    from oracle import modules_ref as mr
    B, S, T, F, E = 2, 3, 4, 129, 50
    energies = torch.tensor([[12.0, -12.0, 0.7], [100.0, -30.0, -3.0]])              # [B,S] real part; imag = -real/2
    # embedding == 1 everywhere: Linear weight 0, bias 20 -> tanh(20) == 1.0f ; <1, q> = sum(q)
    q = torch.zeros(B, S, 2 * E)
    q[:, :, :E] = (energies / E).unsqueeze(-1)
    q[:, :, E:] = (-0.5 * energies / E).unsqueeze(-1)
    e_re = q[:, :, :E].sum(-1)
    e_im = q[:, :, E:].sum(-1)
    k, c = d.config.cRM_k, d.config.cRM_C
    with torch.no_grad():
        want = torch.stack([k * torch.tanh(e_re), k * torch.tanh(e_im)], -1)         # [B,S,2]
        want = -1 / c * torch.log((k - want) / (k + want))
    assert torch.isinf(want[0, 0, 0]) and want[0, 0, 0] > 0 and torch.isinf(want[0, 1, 0]) and want[0, 1, 0] < 0
    assert torch.isinf(want[1, 0, 0]) and torch.isfinite(want[0, 2]).all() and torch.isfinite(want[1, 2]).all()
    old = d.config.is_ComlexMask
    d.config.is_ComlexMask = 1
    try:
        h = torch.zeros(B, T, 600, device=cuda)
        w = torch.zeros(F * E, 600, device=cuda)
        bias = torch.full((F * E,), 20.0, device=cuda)
        fused = d.emb_attn_mask(h, w, bias, q.to(cuda), F, E, complex_mask=True, decompress=True).cpu()   # [B,S,T,F,2]
        emb = torch.ones(B * S, T, F, E, device=cuda)
        att = d.ATTENTION(E, 'dot').to(cuda)
        unfused = d.crm_decompress(att(emb, q.view(-1, 2 * E).to(cuda))).view(B, S, T, F, 2).cpu()
    finally:
        d.config.is_ComlexMask = old
    for got in (fused, unfused):
        for b in range(B):
            for s in range(S):
                for ch in range(2):
                    wv = want[b, s, ch]
                    g = got[b, s, :, :, ch]
                    if torch.isinf(wv):
                        assert torch.isinf(g).all() and (torch.sign(g) == torch.sign(wv)).all(), (b, s, ch, g.flatten()[:3], wv)
                    else:
                        assert torch.isfinite(g).all()
                        tol = 1e-3 if abs(wv.item()) < 60.0 else 1e-2          # conditioning of the decompression
                        assert ((g - wv).abs() / max(abs(wv.item()), 1.0 / c)).max().item() < tol, (b, s, ch, g.flatten()[:3], wv)
    # and the masked reconstruction propagates them the way numpy does: inf * X -> inf/nan samples, never a crash
    spec = torch.randn(B, T, F, 2, device=cuda)
    wav = d.mask_istft(fused.to(cuda), spec, 128)
    assert not torch.isfinite(wav[0, 0]).all().item() and torch.isfinite(wav[:, 2]).all().item()

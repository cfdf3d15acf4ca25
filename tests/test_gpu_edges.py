"""Edge cases of the CUDA path (ragged / empty / maximum shapes), each against torch on the CPU oracle side."""
import numpy as np
import pytest
import torch

from tests.util import build_pair, rel_err

pytestmark = pytest.mark.gpu


def _rnn_pair(cell, layers, H, T, seed=1):
    import dl4ss_b200 as d
    torch.manual_seed(seed)
    rnn = {'lstm': torch.nn.LSTM, 'gru': torch.nn.GRU}[cell](129, H, layers, batch_first=True, bidirectional=True)
    d.config.HIDDEN_UNITS, d.config.NUM_LAYERS = H, layers
    ours = d.MIX_SPEECH(129, T, cell=cell, num_layers=layers).cuda()
    ours.layer.load_state_dict(rnn.state_dict())
    return rnn, ours


@pytest.mark.parametrize('cell,H,B,T', [('lstm', 100, 5, 9), ('gru', 200, 33, 7), ('lstm', 320, 4, 6), ('lstm', 300, 257, 4),
                                         ('gru', 300, 65, 3), ('lstm', 300, 1, 1), ('lstm', 60, 3, 5), ('gru', 90, 2, 5),
                                         ('lstm', 296, 3, 5), ('gru', 316, 2, 4), ('lstm', 4, 2, 3), ('lstm', 256, 34, 3)])
def test_recurrent_shapes(cuda, cell, H, B, T):
    """Every supported hidden size of the tcgen05 kernel (multiples of 20 up to 320), batch sizes that leave
    partial tiles, need a second tile per CTA, or a second launch (> 256), T = 1, and a size that falls back to the
    fp32 CUDA-core kernel (H = 90: a multiple of 10, not of 20).  The h exchange through the output planes: H % 8 == 4
    (reverse window shifted by four columns: 100, 300, 60, 316 -- where H + 4 fills the last k-chunk exactly -- and 4),
    H % 16 == 8 (k-step overhang into the other direction's columns: 200, 296), H a multiple of 64 (exact windows)."""
    import dl4ss_b200 as d
    try:
        rnn, ours = _rnn_pair(cell, 1, H, T)
        x = torch.rand(B, T, 129)
        with torch.no_grad():
            y_ref, _ = rnn(x)
            y = ours.encode(x.cuda()).cpu()
        assert (y - y_ref).abs().max().item() < 2e-5
    finally:
        d.config.HIDDEN_UNITS, d.config.NUM_LAYERS = 300, 2


def test_empty_batches(cuda):
    import dl4ss_b200 as d
    f, c = d.stft_features(torch.zeros(0, 4000, device=cuda), 256, 128)
    assert tuple(f.shape) == (0, 32, 129) and tuple(c.shape) == (0, 32, 129, 2)
    w = d.mask_istft(torch.zeros(0, 2, 32, 129, device=cuda), c, 128)
    assert tuple(w.shape) == (0, 2, 128 * 31)
    assert tuple(d.linear_fwd(torch.zeros(0, 7, device=cuda), torch.zeros(5, 7, device=cuda)).shape) == (0, 5)


@pytest.mark.parametrize('S,cplx', [(1, False), (4, False), (4, True), (1, True)])
def test_speaker_counts(cuda, S, cplx):
    """S = 1 and the largest S the fused epilogue keeps in registers (4; cRM: 8 energies)."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, T = 2, 9
    ref, ours = build_pair('gru' if cplx else 'lstm', 1, 129, T, cplx)
    if cplx:
        with torch.no_grad():
            ref['emb'].layer.weight.mul_(0.1)
            ours['emb'].layer.weight.mul_(0.1)
    torch.manual_seed(2)
    feas = torch.rand(B, T, 129)
    mag = torch.randn(B, T, 129, 2)
    idx = np.sort(np.random.RandomState(S).choice(101, (B, S)), axis=1)
    with torch.no_grad():
        r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, idx, mag)
    m = d.Separator(ours['mix'], ours['emb'], ours['att'], ours['adj']).masks(feas.cuda(), idx).cpu()
    scale = torch.clamp(r['masks'].abs(), min=10.0) if cplx else 1.0
    assert ((m - r['masks']).abs() / scale).max().item() < 1e-4
    d.config.is_ComlexMask = False


@pytest.mark.parametrize('hop,L', [(256, 4096), (64, 1000), (128, 129), (128, 40001)])
def test_stft_istft_edge_lengths(cuda, hop, L):
    """hop = n_fft (no overlap), the shortest signal reflect padding allows, and L not a multiple of hop."""
    import dl4ss_b200 as d
    from oracle import stft_ref as sr
    rng = np.random.RandomState(L)
    wav = rng.standard_normal((2, L))
    ref = np.stack([sr.stft_ref(w, 256, hop).T for w in wav])
    _, c = d.stft_features(torch.from_numpy(wav).cuda(), 256, hop, 'hann', None)
    got = torch.view_as_complex(c).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()
    T = ref.shape[1]
    if T > 1:
        back_ref = np.stack([sr.istft_ref(r.T, hop) for r in ref])
        back = d.mask_istft(None, c.view(2, 1, T, 129, 2), hop).cpu().numpy()[:, 0]
        # where the window sum-square is tiny (hop = n_fft: frame edges of the Hann window) librosa's division
        # amplifies fp32 round-off by 1/w; compare where the envelope is well conditioned
        wss = sr.window_sumsquare('hann', T, 256, hop)[128:128 + back.shape[1]]
        ok = wss > 1e-2
        assert np.abs(back - back_ref)[:, ok].max() < 1e-4 * np.abs(back_ref).max()


def test_errors_are_loud(cuda):
    import dl4ss_b200 as d
    rnn, ours = _rnn_pair('lstm', 1, 322, 4)              # no kernel has a 322-unit decomposition
    try:
        with pytest.raises(RuntimeError, match='multiple'):
            ours.encode(torch.zeros(2, 4, 129, device=cuda))
    finally:
        d.config.HIDDEN_UNITS, d.config.NUM_LAYERS = 300, 2
    with pytest.raises(RuntimeError):
        d.stft_features(torch.zeros(1, 100, device=cuda), 256, 128)          # L <= n_fft/2
    with pytest.raises(RuntimeError):
        d.stft_features(torch.zeros(1, 4000, device=cuda), 512, 128)         # only the 256-point transform
    with pytest.raises(RuntimeError):
        d.mask_istft(torch.zeros(1, 2, 5, 129, device=cuda), torch.zeros(1, 6, 129, 2, device=cuda), 128)
    with pytest.raises(RuntimeError):
        d.linear_fwd(torch.zeros(2, 3, device=cuda).t(), torch.zeros(4, 2, device=cuda))   # non-contiguous


def test_premix_matches_reference_preprocessing(cuda):
    """a1: crop / -mean / /max|.| / zero-pad / dB gain / sum, against the float64 oracle."""
    import dl4ss_b200 as d
    from oracle import stft_ref as sr
    rng = np.random.RandomState(0)
    B, S, L = 3, 3, 9000
    lengths = rng.randint(2000, L + 1, (B, S)); lengths[0, 0] = L
    gains = rng.uniform(-2.5, 2.5, (B, S))
    raw = np.zeros((B, S, L), np.float32)
    for b in range(B):
        for s in range(S):
            raw[b, s, :lengths[b, s]] = (rng.standard_normal(lengths[b, s]) * rng.uniform(0.1, 3) + rng.uniform(-1, 1)).astype(np.float32)
    ref = np.array([[sr.preprocess_source(raw[b, s, :lengths[b, s]].astype(np.float64), L, gains[b, s]) for s in range(S)] for b in range(B)])
    out = d.premix(torch.from_numpy(raw).cuda(), torch.from_numpy(gains), torch.from_numpy(lengths))
    assert np.abs(out['sources'].cpu().numpy() - ref).max() < 2e-6 * np.abs(ref).max()
    assert np.abs(out['mix_wav'].cpu().numpy() - ref.sum(1)).max() < 4e-6 * np.abs(ref).max()
    assert torch.all(out['sources'][1, 1, int(lengths[1, 1]):] == 0)
    # feeds K1 directly: features of the GPU mixture == oracle features of the oracle mixture
    batch = d.prepare_batch(out['mix_wav'], sources=out['sources'])
    f_ref = np.stack([sr.features_ref(m, 256, 128)['mix_feas'] for m in ref.sum(1)])
    assert rel_err(batch['mix_feas'].cpu().numpy(), f_ref) < 1e-4
    assert tuple(batch['multi_spk_fea'].shape) == (B, S, 1 + L // 128, 129)


def test_classifier_and_speaker_selection(cuda):
    """n1: MIX_SPEECH_classifier (BLSTM 3x600 -> mean -> Linear -> sigmoid) against the oracle, then the
    reference's eval flow: top_k_mask of its output -> speaker ids -> separation (EvalVer.py:424-470)."""
    import copy
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, T = 3, 11
    ref, ours = build_pair('lstm', 1, 129, T, False)
    torch.manual_seed(7)
    cls_ref = mr.MIX_SPEECH_classifier(ref['cfg'], 129, T, 101)
    cls = d.MIX_SPEECH_classifier(129, T, 101).cuda()
    assert list(cls.state_dict().keys()) == list(cls_ref.state_dict().keys())
    cls.load_state_dict(copy.deepcopy(cls_ref.state_dict()))
    feas = torch.rand(B, T, 129) * 2
    with torch.no_grad():
        p_ref = cls_ref(feas)
        p = cls(feas.cuda())
    assert tuple(p.shape) == (B, 101)
    assert (p.cpu() - p_ref).abs().max().item() < 2e-5
    # test-mode selection of the reference: alpha = -0.5, top_k = 2 -> exactly two speakers per utterance
    mask = d.top_k_mask(p, -0.5, 2)
    assert mask.is_cuda and torch.equal(mask.cpu(), mr.top_k_mask(p_ref, -0.5, 2))
    idx = torch.nonzero(mask)[:, 1].view(B, 2)
    with torch.no_grad():
        r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, idx.cpu().numpy())
    m = d.Separator(ours['mix'], ours['emb'], ours['att'], ours['adj']).masks(feas.cuda(), idx)
    assert (m.cpu() - r['masks']).abs().max().item() < 1e-4


def test_recursive_extract_matches_oracle(cuda):
    """n4: the recursive extract-and-subtract inference (RecuVer.py:32-79,480-494) on the device, whole
    batch at once, against the oracle's restatement of the reference loop: same speaker named at every step,
    predicted spectrograms within 1e-4 of the mixture scale, and the residual shrinks step by step."""
    import copy
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, T, steps = 4, 13, 3
    ref, ours = build_pair('lstm', 2, 129, T, False)
    torch.manual_seed(11)
    cls_ref = mr.MIX_SPEECH_classifier(ref['cfg'], 129, T, 101)
    with torch.no_grad():
        cls_ref.Linear.weight.mul_(40.0)      # a decisive head: top-1 margins >= 2e-3, far above the 2e-5 fp32 path difference
    cls = d.MIX_SPEECH_classifier(129, T, 101).cuda()
    cls.load_state_dict(copy.deepcopy(cls_ref.state_dict()))
    feas = torch.rand(B, T, 129) * 2
    with torch.no_grad():
        want, spk_ref = mr.recursive_extract_ref(ref['cfg'], cls_ref, ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, steps)
    sep = d.Separator(ours['mix'], ours['emb'], ours['att'], ours['adj'])
    got, spk = sep.recursive_extract(feas.cuda(), cls, steps)
    assert tuple(got.shape) == (B, steps, T, 129) and tuple(spk.shape) == (B, steps)
    assert np.array_equal(spk.cpu().numpy(), spk_ref)
    assert (got.cpu() - want).abs().max().item() < 1e-4 * 2.0
    resid = feas.cuda().unsqueeze(1) - got.cumsum(1)
    assert (resid[:, -1].abs().sum() < resid[:, 0].abs().sum()).item()


@pytest.mark.parametrize('cell,B,T,H', [('lstm', 5, 7, 600), ('gru', 37, 6, 600), ('lstm', 70, 5, 300), ('gru', 3, 9, 40),
                                         ('lstm', 16, 1, 600)])
def test_rnn_mma_layer_equals_fp32_layer(cuda, cell, B, T, H):
    """dl4ss_rnn_layer_mma_fwd (warp-level tensor cores, bf16x3; the H = 600 classifier path) against the fp32
    recurrent kernel on the same hoisted projection: outputs, saved gates and cells; partial and multiple tiles."""
    import ctypes
    from dl4ss_b200 import _lib as L
    lib = L.load()
    G = 4 if cell == 'lstm' else 3
    c = L.CELL_LSTM if cell == 'lstm' else L.CELL_GRU
    assert lib.dl4ss_rnn_mma_supported(H, c) == 1
    g = torch.Generator().manual_seed(B * 131 + H)
    xproj = torch.randn(B, T, 2, G * H, generator=g).to(cuda)
    whh = (torch.randn(2, G * H, H, generator=g) / H ** 0.5).to(cuda)
    bhn = torch.randn(2, H, generator=g).to(cuda) if G == 3 else None

    def run(fn, wsb):
        y = torch.full((B, T, 2 * H), float('nan'), device=cuda)
        gates = torch.full((B, T, 2, G * H), float('nan'), device=cuda)
        cells = torch.full((B, T, 2, H), float('nan'), device=cuda)
        need = wsb(B, T, H, c)
        ws = torch.empty(need, device=cuda, dtype=torch.uint8)
        rc = fn(c, L.ptr(xproj), L.ptr(whh), L.ptr(bhn), L.ptr(y), B, T, H, L.ptr(gates), L.ptr(cells),
                ctypes.c_void_p(ws.data_ptr()), need, L.stream())
        L.check(rc, 'rnn layer')
        torch.cuda.synchronize()
        return y, gates, cells

    y0, g0, c0 = run(lib.dl4ss_rnn_layer_fwd, lib.dl4ss_rnn_workspace_bytes)
    y1, g1, c1 = run(lib.dl4ss_rnn_layer_mma_fwd, lib.dl4ss_rnn_mma_workspace_bytes)
    assert (y1 - y0).abs().max().item() < 2e-5
    assert (g1 - g0).abs().max().item() < 2e-5
    assert (c1 - c0).abs().max().item() < 2e-5 * max(1.0, c0.abs().max().item())


def test_rnn_mma_unsupported_is_loud(cuda):
    from dl4ss_b200 import _lib as L
    lib = L.load()
    assert lib.dl4ss_rnn_mma_supported(601, L.CELL_LSTM) == 0
    assert lib.dl4ss_rnn_mma_supported(1000, L.CELL_LSTM) == 0
    x = torch.zeros(16, device=cuda)
    rc = lib.dl4ss_rnn_layer_mma_fwd(L.CELL_LSTM, L.ptr(x), L.ptr(x), None, L.ptr(x), 1, 1, 601, None, None, None, 0, L.stream())
    assert rc != 0 and b'unsupported' in lib.dl4ss_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize('bs,topk,T,F', [(2, 2, 313, 129), (1, 3, 61, 129), (3, 1, 313, 129)])
def test_discriminator_matches_oracle(cuda, bs, topk, T, F):
    """n4: Discriminator forward through the C ABI (direct 3x3 stride-2 convolutions + the 36480-wide score layer) against
    the oracle's transliteration (TDAA_beta/main_run_sstune_EvalVer.py:328-346) with the same weights, and the
    least-squares GAN terms of the training loop (:643-652, :670-671)."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    torch.manual_seed(11)
    flat = d.Discriminator.flat_features(T, F)
    ref = mr.Discriminator(flat)
    ours = d.Discriminator(flat)
    ours.load_state_dict(ref.state_dict())          # same keys: cnn / cnn1 / cnn2 / final
    ours = ours.to(cuda)
    true_map = torch.rand(bs, topk, T, F) * 2.0
    false_map = torch.rand(bs, topk, T, F)
    with torch.no_grad():
        r_true, r_false = ref(true_map), ref(false_map)
        o_true, o_false = ours(true_map.to(cuda)), ours(false_map.to(cuda))
    assert o_true.shape == (bs * topk, 1)
    assert (o_true.cpu() - r_true).abs().max().item() < 1e-5
    assert (o_false.cpu() - r_false).abs().max().item() < 1e-5
    lr, lo = mr.gan_loss_terms_ref(r_true, r_false), d.gan_loss_terms(o_true, o_false)
    for k in lr:
        assert abs(float(lo[k]) - float(lr[k])) < 1e-5, k
    # the intermediate activations as well: first two layers against torch's convolution
    import torch.nn.functional as Fn
    x = true_map.view(bs * topk, 1, T, F)
    a1 = Fn.relu(ref.cnn(x))
    lib = d.load_library()
    from dl4ss_b200 import _lib
    y1 = torch.empty(a1.shape, device=cuda)
    rc = lib.dl4ss_conv3x3s2_relu_fwd(_lib.ptr(x.to(cuda).contiguous()), _lib.ptr(ours.cnn.weight.detach()),
                                      _lib.ptr(ours.cnn.bias.detach()), _lib.ptr(y1), bs * topk, 1, T, F, 64, _lib.stream())
    assert rc == 0
    assert (y1.cpu() - a1.detach()).abs().max().item() < 1e-5
    a2 = Fn.relu(ref.cnn1(a1))
    y2 = torch.empty(a2.shape, device=cuda)
    rc = lib.dl4ss_conv3x3s2_relu_fwd(_lib.ptr(y1), _lib.ptr(ours.cnn1.weight.detach()), _lib.ptr(ours.cnn1.bias.detach()),
                                      _lib.ptr(y2), bs * topk, 64, a1.shape[2], a1.shape[3], 64, _lib.stream())
    assert rc == 0
    assert (y2.cpu() - a2.detach()).abs().max().item() < 2e-5 * max(1.0, a2.abs().max().item())
    with pytest.raises(RuntimeError):               # an unsupported layer shape is an error, not a fallback
        rc = lib.dl4ss_conv3x3s2_relu_fwd(_lib.ptr(y1), _lib.ptr(ours.cnn1.weight.detach()), None, _lib.ptr(y2),
                                          bs * topk, 24, 10, 10, 64, _lib.stream())
        _lib.check(rc, 'dl4ss_conv3x3s2_relu_fwd')

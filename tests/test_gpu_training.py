"""GPU parity of the training step: loss and every parameter gradient against torch autograd on the
CPU oracle (the reference's `loss.backward()`, TDAA_beta/main_run_sstune_EvalVer.py:673)."""
import numpy as np
import pytest
import torch

from tests.util import build_pair

pytestmark = pytest.mark.gpu


def _grads(mods):
    out = {}
    for name, m in mods.items():
        if m is None or name == 'cfg':
            continue
        for k, p in m.named_parameters():
            out[name + '.' + k] = None if p.grad is None else p.grad.detach().cpu().double()
    return out


@pytest.mark.parametrize('cell,layers,cplx,S,B,T', [('lstm', 2, False, 2, 3, 19), ('gru', 2, True, 3, 2, 14),
                                                     ('lstm', 1, False, 2, 2, 7), ('gru', 1, False, 2, 4, 11),
                                                     # BASELINE configs[3] depth and length: 4 layers, the full 313-step chain,
                                                     # more utterances than one 16-row tile
                                                     ('lstm', 4, False, 2, 18, 313)])
def test_gradients_match_autograd(cuda, cell, layers, cplx, S, B, T):
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    ref, ours = build_pair(cell, layers, 129, T, cplx)
    torch.manual_seed(21)
    feas = torch.rand(B, T, 129) * 2
    mag = torch.randn(B, T, 129, 2)
    y = torch.rand(B, S, T, 129, 2) if cplx else torch.rand(B, S, T, 129)
    if cplx:                      # keep the cRM energies small: the reference's log/tanh chain is ill-conditioned
        with torch.no_grad():
            ref['emb'].layer.weight.mul_(0.05)
            ours['emb'].layer.weight.mul_(0.05)
    idx = np.sort(np.random.RandomState(4).choice(101, (B, S)), axis=1)
    r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, idx, mag)
    loss = mr.loss_ref(ref['cfg'], r, y)[0]
    loss.backward()
    g_ref = _grads(ref)

    step = d.TrainStep(ours['mix'], ours['emb'], ours['att'], ours['adj'])
    l, _, _ = step.loss_and_grads(feas.to(cuda), idx, y.to(cuda).contiguous(), mag.to(cuda).contiguous())
    assert abs(l.item() - loss.item()) < 1e-5 * abs(loss.item()) + 1e-8
    g = _grads(ours)
    checked = 0
    for k, gr in g_ref.items():
        if k.startswith('att.'):              # Linear_1/2/3 are unused by the dot attention: no gradient
            continue
        assert g[k] is not None, k
        scale = max(gr.abs().max().item(), 1e-12)
        err = (g[k] - gr).abs().max().item() / scale
        assert err < 2e-3, (k, err)
        checked += 1
    assert checked >= 8 * layers + 3


def test_train_step_reduces_loss(cuda):
    """A few Adam steps through the public TrainStep.step lower the loss (the reference's lr 2e-4 x 10)."""
    import dl4ss_b200 as d
    _, ours = build_pair('lstm', 2, 129, 16, False)
    step = d.TrainStep(ours['mix'], ours['emb'], ours['att'], ours['adj'])
    opt = torch.optim.Adam([{'params': step.parameters()}], lr=2e-3)
    torch.manual_seed(3)
    feas = (torch.rand(4, 16, 129) * 2).to(cuda)
    y = (torch.rand(4, 2, 16, 129) * feas.cpu().unsqueeze(1)).to(cuda).contiguous()
    idx = np.array([[1, 5], [2, 9], [0, 7], [3, 4]])
    losses = [step.step(opt, feas, idx, y)[0].item() for _ in range(8)]
    assert losses[-1] < losses[0]
    # the gradients live in ONE persistent bucket (views, no copies), cut into head + one segment per layer
    b = step.bucket()
    assert len(b.ranges) == 1 + 2 and b.ranges[-1][1] == b.flat.numel() == sum(p.numel() for p in step.parameters())
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b.views))
    assert sorted(id(p) for p in b.params) == sorted(id(p) for p in step.parameters())
    assert step.reduced_bytes == 0                            # one process: nothing to exchange


def _bptt_step_chain(lib, L, cell, dy, whh, gates, cells, y, B, T, H, G):
    """The per-step form (library GEMM + dl4ss_rnn_bwd_step): the yard-stick for the persistent kernel."""
    dev = dy.device
    dgx = torch.zeros(B, T, 2, G * H, device=dev)
    dgh = torch.zeros(B, T, 2, G * H, device=dev) if G == 3 else None
    carry = torch.zeros(2, B, H, device=dev)
    dg_cur = torch.zeros(2, B, G * H, device=dev)
    dh_rec = torch.zeros(2, B, H, device=dev)
    for s in range(T):
        if s > 0:
            torch.bmm(dg_cur, whh, out=dh_rec)
        rc = lib.dl4ss_rnn_bwd_step(cell, s, L.ptr(dy), L.ptr(dh_rec), L.ptr(gates), L.ptr(cells), L.ptr(y),
                                    L.ptr(carry), L.ptr(dgx), L.ptr(dgh), L.ptr(dg_cur), B, T, H, L.stream())
        L.check(rc, 'dl4ss_rnn_bwd_step')
    return dgx, dgh


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('cell,B,T,H', [('lstm', 5, 9, 300), ('gru', 37, 6, 300), ('lstm', 70, 4, 300),
                                         ('gru', 16, 1, 300), ('lstm', 3, 5, 40)])
def test_persistent_bptt_equals_step_chain(cuda, cell, B, T, H, tc):
    """dl4ss_rnn_layer_bwd / dl4ss_rnn_layer_bwd_tc (one persistent kernel; fp32 FMA and bf16x3 mma forms) against the
    per-step chain on the same saved activations: partial tiles, several tiles, more utterances than one launch holds
    (B=70), T=1, a small H."""
    import ctypes
    from dl4ss_b200 import _lib as L
    lib = L.load()
    G = 4 if cell == 'lstm' else 3
    c = L.CELL_LSTM if cell == 'lstm' else L.CELL_GRU
    assert lib.dl4ss_rnn_bwd_supported(H, c) == 1
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(B, T, 2 * H, generator=g).to(cuda)
    whh = (torch.randn(2, G * H, H, generator=g) / H ** 0.5).to(cuda)
    gates = torch.rand(B, T, 2, G * H, generator=g)
    if cell == 'lstm':
        gates[..., 2 * H:3 * H] = gates[..., 2 * H:3 * H] * 2 - 1        # g gate is a tanh
    else:
        gates[..., 2 * H:] = gates[..., 2 * H:] * 2 - 1                  # n gate is a tanh
    gates = gates.to(cuda)
    cells = torch.randn(B, T, 2, H, generator=g).to(cuda)
    y = (torch.rand(B, T, 2 * H, generator=g) * 2 - 1).to(cuda)
    ref_x, ref_h = _bptt_step_chain(lib, L, c, dy, whh, gates, cells, y, B, T, H, G)
    dgx = torch.full((B, T, 2, G * H), float('nan'), device=cuda)
    dgh = torch.full((B, T, 2, G * H), float('nan'), device=cuda) if G == 3 else None
    need = lib.dl4ss_rnn_bwd_workspace_bytes(B, T, H, c)
    ws = torch.empty(need, device=cuda, dtype=torch.uint8)
    if tc:
        assert lib.dl4ss_rnn_bwd_tc_supported(H, c) == 1
        xp = torch.zeros(lib.dl4ss_rnn_bwd_tc_xplanes_bytes(B, T, H, c), device=cuda, dtype=torch.uint8)
        rc = lib.dl4ss_rnn_layer_bwd_tc(c, L.ptr(dy), L.ptr(whh), L.ptr(gates), L.ptr(cells), L.ptr(y), L.ptr(dgx),
                                        L.ptr(dgh), ctypes.c_void_p(xp.data_ptr()), B, T, H,
                                        ctypes.c_void_p(ws.data_ptr()), need, L.stream())
    else:
        rc = lib.dl4ss_rnn_layer_bwd(c, L.ptr(dy), L.ptr(whh), L.ptr(gates), L.ptr(cells), L.ptr(y), L.ptr(dgx),
                                     L.ptr(dgh), B, T, H, ctypes.c_void_p(ws.data_ptr()), need, L.stream())
    L.check(rc, 'dl4ss_rnn_layer_bwd')
    torch.cuda.synchronize()
    scale = ref_x.abs().max().item()
    tol = 1e-4 if tc else 2e-5          # bf16x3 drops the lo*lo products (2^-16 relative per product)
    assert (dgx - ref_x).abs().max().item() < tol * scale
    if G == 3:
        assert (dgh - ref_h).abs().max().item() < tol * scale


def test_persistent_bptt_unsupported_is_loud(cuda):
    from dl4ss_b200 import _lib as L
    lib = L.load()
    assert lib.dl4ss_rnn_bwd_supported(301, L.CELL_LSTM) == 0
    assert lib.dl4ss_rnn_bwd_supported(600, L.CELL_LSTM) == 0          # 20 x 2400 fp32 W slice + tile do not fit
    x = torch.zeros(16, device=cuda)
    rc = lib.dl4ss_rnn_layer_bwd(L.CELL_LSTM, L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), None, L.ptr(x), None, 1, 1, 301,
                                 None, 0, L.stream())
    assert rc != 0 and b'unsupported' in lib.dl4ss_last_error()


@pytest.mark.parametrize('R,Ca,Cb', [(5000, 300, 129), (777, 1200, 300), (64, 40, 7), (20032, 130, 600)])
def test_matmul_tn_split_k(cuda, R, Ca, Cb):
    """dW-shaped contraction a^T b over many rows and few output tiles: the split-K launch (partial tiles summed with
    atomics) against float64, and with a bias against the single-pass launch."""
    from dl4ss_b200 import modules as M
    g = torch.Generator().manual_seed(R)
    a = torch.randn(R, Ca, generator=g)
    b = torch.randn(R, Cb, generator=g)
    ref = a.double().t() @ b.double()
    out = M.matmul_tn(a.to(cuda), b.to(cuda))
    torch.cuda.synchronize()
    assert out.shape == (Ca, Cb)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()
    bias = torch.randn(Cb, generator=g).to(cuda)
    ap, bp = M.split_bf16_t(a.to(cuda)), M.split_bf16_t(b.to(cuda))
    one = M.linear_tc(ap, bp, bias, Ca, Cb, R)
    many = M.linear_tc(ap, bp, bias, Ca, Cb, R, split_k=True)
    torch.cuda.synchronize()
    assert (one - many).abs().max().item() < 1e-4 * ref.abs().max().item()      # single-pass fp32 accumulation over all of K is the looser of the two


@pytest.mark.parametrize('S,cplx', [(2, False), (3, False), (4, False), (3, True)])
def test_pit_mask_loss_matches_oracle(cuda, S, cplx):
    """dl4ss_mask_pair_loss_fwd + permutation search against the oracle's brute-force PIT (oracle/modules_ref.py
    pit_mse_ref): loss value and the chosen assignment, with the targets of half the utterances shuffled."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, T, F = 6, 11, 129
    g = torch.Generator().manual_seed(S + 10 * cplx)
    if cplx:
        masks = torch.randn(B, S, T, F, 2, generator=g)
        mix = torch.randn(B, T, F, 2, generator=g)
        pr = masks[..., 0] * mix[:, None, ..., 0] - masks[..., 1] * mix[:, None, ..., 1]
        pi = masks[..., 0] * mix[:, None, ..., 1] + masks[..., 1] * mix[:, None, ..., 0]
        pred = torch.stack([pr, pi], -1)
    else:
        masks = torch.rand(B, S, T, F, generator=g)
        mix = torch.rand(B, T, F, generator=g) * 2
        pred = masks * mix[:, None]
    target = pred + 0.05 * torch.randn(pred.shape, generator=g)
    true_perm = torch.stack([torch.randperm(S, generator=g) if b % 2 else torch.arange(S) for b in range(B)])
    shuffled = torch.empty_like(target)
    for b in range(B):
        shuffled[b, true_perm[b]] = target[b]            # target s now sits at position true_perm[b][s]
    ref_loss, ref_perm = mr.pit_mse_ref(pred, shuffled)
    loss, perms = d.pit_mask_loss(masks.to(cuda).contiguous(), mix.to(cuda).contiguous(), shuffled.to(cuda).contiguous())
    scale = 2.0 if cplx else 1.0                           # cRM: MSE(Re) + MSE(Im) = 2 x the mean over the stacked pair
    assert abs(loss.item() - scale * ref_loss.item()) < 1e-6 * abs(ref_loss.item()) + 1e-9
    assert torch.equal(perms.cpu(), ref_perm)
    assert torch.equal(perms.cpu(), true_perm)


def test_pit_training_gradients(cuda):
    """TrainStep.loss_and_grads(pit=True): the gradients are those of the reference loss with the targets re-ordered to
    the best assignment (autograd on the oracle with the oracle's own PIT assignment)."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    B, S, T = 4, 2, 13
    ref, ours = build_pair('lstm', 1, 129, T, False)
    torch.manual_seed(9)
    feas = torch.rand(B, T, 129) * 2
    idx = np.sort(np.random.RandomState(4).choice(101, (B, S)), axis=1)
    r = mr.forward_ref(ref['cfg'], ref['mix'], ref['emb'], ref['att'], ref['adj'], feas, idx, None)
    # targets = the current predictions pushed apart (source 0 up, source 1 down) plus noise, in swapped order for
    # utterances 1 and 3: their best assignment is the swap
    y = r['predict'].detach() * torch.tensor([1.5, 0.5]).view(1, S, 1, 1) + 0.02 * torch.rand(B, S, T, 129)
    y[1] = y[1].flip(0)
    y[3] = y[3].flip(0)
    _, perm = mr.pit_mse_ref(r['predict'].detach(), y)
    assert perm.tolist() == [[0, 1], [1, 0], [0, 1], [1, 0]]
    y_perm = y[torch.arange(B)[:, None], perm]
    loss = mr.loss_ref(ref['cfg'], r, y_perm)[0]
    loss.backward()
    g_ref = _grads(ref)
    step = d.TrainStep(ours['mix'], ours['emb'], ours['att'], ours['adj'])
    l, _, _ = step.loss_and_grads(feas.to(cuda), idx, y.to(cuda).contiguous(), None, pit=True)
    assert abs(l.item() - loss.item()) < 1e-5 * abs(loss.item()) + 1e-8
    g = _grads(ours)
    for k, gr in g_ref.items():
        if k.startswith('att.'):
            continue
        scale = max(gr.abs().max().item(), 1e-12)
        assert (g[k] - gr).abs().max().item() / scale < 2e-3, k


@pytest.mark.gpu
@pytest.mark.parametrize('B,T,Ca,Cb,col0_a,M,col0_b,N,sa,sb', [(3, 70, 200, 129, 0, 200, 0, 129, 0, 0),
                                                               (5, 313, 2400, 600, 1200, 1200, 300, 300, 0, 1),
                                                               (4, 64, 320, 600, 0, 300, 0, 300, 0, -1),
                                                               (2, 130, 136, 72, 8, 100, 3, 60, 1, 0),
                                                               (2, 100, 1208, 608, 8, 1200, 0, 601, -1, 0)])   # swapped operands, transposed epilogue
def test_linear_tc_tn_matches_float64(cuda, B, T, Ca, Cb, col0_a, M, col0_b, N, sa, sb):
    """MN-major split-K product (dl4ss_linear_tc_tn_splitk_fwd): sum over (b,t) of A[b,t+sa,col0_a+m] * B[b,t+sb,col0_b+n]
    from row-major bf16 hi/lo planes, against float64 -- column windows, frame shifts (out-of-range frames are zeros),
    T not a multiple of the 64-frame k-block, pitches that are not multiples of 64."""
    import dl4ss_b200 as d
    from dl4ss_b200 import modules as Mo
    g = torch.Generator().manual_seed(B * 1000 + T)
    a = torch.randn(B, T, Ca, generator=g)
    b = torch.randn(B, T, Cb, generator=g)

    def planes(x):            # [2, B*T, ld] with ld = columns rounded up to 8 (not the 64 the K-major path wants)
        C = x.shape[-1]
        ld = (C + 7) // 8 * 8
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        out = torch.zeros(2, x.shape[0] * x.shape[1], ld, dtype=torch.bfloat16)
        out[0, :, :C], out[1, :, :C] = hi.view(-1, C), lo.view(-1, C)
        return out.to(cuda)

    got = Mo.linear_tc_tn(planes(a), col0_a, M, sa, planes(b), col0_b, N, sb, B, T, wa=Ca, wb=Cb).cpu().double()

    def shifted(x, s):
        y = torch.zeros_like(x)
        if s == 0:
            y = x.clone()
        elif s > 0:
            y[:, :-s] = x[:, s:]
        else:
            y[:, -s:] = x[:, :s]
        return y

    want = torch.einsum('btm,btn->mn', shifted(a, sa)[:, :, col0_a:col0_a + M].double(), shifted(b, sb)[:, :, col0_b:col0_b + N].double())
    err = (got - want).abs().max().item()
    assert err < 3e-5 * want.abs().max().item(), err


@pytest.mark.gpu
@pytest.mark.parametrize('crm_c', [0.1, 0.0])
def test_attn_dot_bwd_planes_crm(cuda, crm_c):
    """cRM energies (two per speaker): the planes form equals the bf16 split of the fp32 form, both derivative branches
    (crm_c > 0: the -1/C log decompression; crm_c == 0: the compressed K tanh mask), S = 3, rows not a multiple of a pass."""
    from dl4ss_b200 import _lib
    lib = _lib.load()
    B, S, T, F, E = 2, 3, 5, 129, 50
    g = torch.Generator(device='cuda').manual_seed(5)
    emb = torch.tanh(torch.randn(B, T * F, E, device=cuda, generator=g))
    q = torch.randn(B, S, 2 * E, device=cuda, generator=g) * 0.3
    mask = torch.rand(B, S, T * F, 2, device=cuda, generator=g) * 4 - 2
    dmask = torch.randn(B, S, T * F, 2, device=cuda, generator=g)
    dz = torch.empty_like(emb)
    dq = torch.empty_like(q)
    assert lib.dl4ss_attn_dot_bwd(_lib.ptr(emb), _lib.ptr(q), _lib.ptr(mask), _lib.ptr(dmask), B, S, T * F, E, _lib.ATT_DOT_CRM,
                                  10.0, crm_c, _lib.ptr(dz), _lib.ptr(dq), _lib.stream()) == 0
    ldp = (F * E + 7) // 8 * 8
    planes = torch.zeros(2, B * T, ldp, device=cuda, dtype=torch.bfloat16)
    dq2 = torch.empty_like(q)
    assert lib.dl4ss_attn_dot_bwd_planes(_lib.ptr(emb), _lib.ptr(q), _lib.ptr(mask), _lib.ptr(dmask), B, S, T, F, E, _lib.ATT_DOT_CRM,
                                         10.0, crm_c, _lib.ptr(planes, torch.bfloat16), ldp, _lib.ptr(dq2), _lib.stream()) == 0
    torch.cuda.synchronize()
    dz2d = dz.view(B * T, F * E)
    hi = dz2d.to(torch.bfloat16)
    lo = (dz2d - hi.float()).to(torch.bfloat16)
    assert torch.equal(planes[0, :, :F * E], hi) and torch.equal(planes[1, :, :F * E], lo)
    assert torch.allclose(dq, dq2, rtol=1e-5, atol=1e-5 * float(dq.abs().max()))
    # float64 restatement of the fp32 form itself
    dev = dmask.double() * (2.0 / crm_c if crm_c > 0 else (10.0 - mask.double() ** 2 / 10.0))      # [B,S,TF,2]
    qd = q.double().view(B, S, 2, E)
    want_dz = torch.einsum('bsrc,bsce->bre', dev, qd) * (1 - emb.double() ** 2)
    want_dq = torch.einsum('bsrc,bre->bsce', dev, emb.double()).reshape(B, S, 2 * E)
    assert (dz.double() - want_dz).abs().max().item() < 1e-5 * want_dz.abs().max().item()
    assert (dq2.double() - want_dq).abs().max().item() < 1e-5 * want_dq.abs().max().item()


def test_attn_dot_bwd_planes_and_lda_projection(cuda):
    """dl4ss_attn_dot_bwd_planes emits exactly the bf16 hi/lo split of what dl4ss_attn_dot_bwd writes in fp32 (same dq), in
    the [2][B*T][ldp] layout; dl4ss_linear_tc_lda_fwd multiplies those planes (row pitch ldp, K = F*E not a multiple of 64)
    with a weight to within bf16x3 accuracy of float64."""
    import dl4ss_b200 as d
    from dl4ss_b200 import _lib, modules as Mo
    lib = _lib.load()
    B, S, T, F, E = 3, 2, 7, 129, 50
    g = torch.Generator(device='cuda').manual_seed(3)
    emb = torch.tanh(torch.randn(B, T * F, E, device=cuda, generator=g))
    q = torch.randn(B, S, E, device=cuda, generator=g) * 0.3
    mask = torch.rand(B, S, T * F, device=cuda, generator=g)
    dmask = torch.randn(B, S, T * F, device=cuda, generator=g)
    dz = torch.empty_like(emb)
    dq = torch.empty_like(q)
    rc = lib.dl4ss_attn_dot_bwd(_lib.ptr(emb), _lib.ptr(q), _lib.ptr(mask), _lib.ptr(dmask), B, S, T * F, E, _lib.ATT_DOT,
                                10.0, 0.1, _lib.ptr(dz), _lib.ptr(dq), _lib.stream())
    assert rc == 0
    ldp = (F * E + 7) // 8 * 8
    planes = torch.zeros(2, B * T, ldp, device=cuda, dtype=torch.bfloat16)
    dq2 = torch.empty_like(q)
    rc = lib.dl4ss_attn_dot_bwd_planes(_lib.ptr(emb), _lib.ptr(q), _lib.ptr(mask), _lib.ptr(dmask), B, S, T, F, E, _lib.ATT_DOT,
                                       10.0, 0.1, _lib.ptr(planes, torch.bfloat16), ldp, _lib.ptr(dq2), _lib.stream())
    assert rc == 0
    torch.cuda.synchronize()
    dz2d = dz.view(B * T, F * E)
    hi = dz2d.to(torch.bfloat16)
    lo = (dz2d - hi.float()).to(torch.bfloat16)
    assert torch.equal(planes[0, :, :F * E], hi) and torch.equal(planes[1, :, :F * E], lo)
    assert float(planes[:, :, F * E:].abs().max()) == 0.0
    assert torch.allclose(dq, dq2, rtol=1e-5, atol=1e-5 * float(dq.abs().max())), (dq - dq2).abs().max()   # summation order differs
    # dh = dz W from the planes as they lie
    K, N = F * E, 600
    w = torch.randn(K, N, device=cuda, generator=g) / K ** 0.5            # [K, N]: dh = dz @ w
    out = torch.empty(B * T, N, device=cuda)
    rc = lib.dl4ss_linear_tc_lda_fwd(_lib.ptr(planes, torch.bfloat16), ldp, _lib.ptr(Mo.split_bf16(w.t().contiguous()), torch.bfloat16),
                                     None, _lib.ptr(out), N, B * T, N, K, _lib.stream())
    assert rc == 0
    want = dz2d.double() @ w.double()
    assert (out.double() - want).abs().max().item() < 3e-5 * max(1e-6, want.abs().max().item())

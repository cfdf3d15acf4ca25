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
                                                     ('lstm', 1, False, 2, 2, 7), ('gru', 1, False, 2, 4, 11)])
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

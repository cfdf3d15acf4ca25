"""Python-3 / torch-CPU transliteration of the reference's separation modules (oracle, test infra).

Behaviour follows, line for line in meaning (not in text):
  MIX_SPEECH        TDAA_beta/main_run_sstune_EvalVer.py:277-303 (LSTM, 4 layers hard-coded)
                    Torch_multi/test_multi_labels_speech.py:212-233 (LSTM, NUM_LAYERS)
                    TDAA_beta/main_run_sstune_cRM_EvalVer.py:340-365 (GRU, NUM_LAYERS)
  ATTENTION         TDAA_beta/main_run_sstune_EvalVer.py:199-242 ; cRM dot ...cRM_EvalVer.py:247-271
  SPEECH_EMBEDDING  TDAA_beta/main_run_sstune_EvalVer.py:348-361 ; cRM (2E) ...cRM_EvalVer.py:390-406
                    old multi-hot form Torch_multi/main_run_multi_selfSS.py:308-328
  ADDJUST           TDAA_beta/main_run_sstune_EvalVer.py:363-377 ; cRM ...cRM_EvalVer.py:408-426
  top_k_mask        TDAA_beta/main_run_sstune_EvalVer.py:390-405
  forward glue      TDAA_beta/main_run_sstune_EvalVer.py:420-497 ; ...cRM_EvalVer.py:498-568
The `.cuda()` calls and `Variable` wrappers of the PyTorch-0.3 original are dropped; B is taken
from the input instead of config.BATCH_SIZE (same value in the reference loops).
The arithmetic is PyTorch's own (nn.LSTM/GRU/Linear/Embedding, baddbmm, sigmoid, tanh,
MSELoss) on CPU fp32 (or fp64 when the modules are .double()).
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

cRM_k = 10.0      # TDAA_beta/main_run_sstune_cRM_EvalVer.py:28
cRM_C = 0.1       # TDAA_beta/main_run_sstune_cRM_EvalVer.py:29


class RefConfig(object):
    """The config globals the modules read (TDAA_beta/config_WSJ0_dB.py:77-153)."""

    def __init__(self, **kw):
        self.HIDDEN_UNITS = 300
        self.NUM_LAYERS = 2
        self.EMBEDDING_SIZE = 50
        self.FRAME_RATE = 8000
        self.FRAME_LENGTH = 256
        self.FRAME_SHIFT = 128
        self.MAX_LEN = 40000
        self.is_ComlexMask = False
        self.is_SelfTune = True
        self.IS_LOG_SPECTRAL = False
        for k, v in kw.items():
            setattr(self, k, v)


class MIX_SPEECH(nn.Module):
    def __init__(self, config, input_fre, mix_speech_len, cell='lstm', num_layers=None):
        super(MIX_SPEECH, self).__init__()
        self.config = config
        self.input_fre = input_fre
        self.mix_speech_len = mix_speech_len
        rnn = {'lstm': nn.LSTM, 'gru': nn.GRU}[cell]
        self.layer = rnn(input_size=input_fre, hidden_size=config.HIDDEN_UNITS,
                         num_layers=num_layers if num_layers is not None else config.NUM_LAYERS,
                         batch_first=True, bidirectional=True)
        self.Linear = nn.Linear(2 * config.HIDDEN_UNITS, self.input_fre * config.EMBEDDING_SIZE)

    def forward(self, x):
        B = x.size(0)
        x, hidden = self.layer(x)
        x = x.contiguous()
        xx = x
        x = x.view(B * self.mix_speech_len, -1)
        out = self.Linear(x)
        out = torch.tanh(out)
        out = out.view(B, self.mix_speech_len, self.input_fre, -1)
        return out, xx


class MIX_SPEECH_classifier(nn.Module):
    """Speaker-presence classifier ("who is talking"): BLSTM 3 x (2*HIDDEN_UNITS) -> mean over T -> Linear -> sigmoid
    (TDAA_beta/main_run_sstune_EvalVer.py:305-326).  Its top-k speakers are the queries of the separation path."""

    def __init__(self, config, input_fre, mix_speech_len, num_labels):
        super(MIX_SPEECH_classifier, self).__init__()
        self.input_fre = input_fre
        self.mix_speech_len = mix_speech_len
        self.layer = nn.LSTM(input_size=input_fre, hidden_size=2 * config.HIDDEN_UNITS, num_layers=3,
                             batch_first=True, bidirectional=True)
        self.Linear = nn.Linear(2 * 2 * config.HIDDEN_UNITS, num_labels)

    def forward(self, x):
        x, hidden = self.layer(x)
        x = x.contiguous()
        x = torch.mean(x, 1)
        return torch.sigmoid(self.Linear(x))


class Discriminator(nn.Module):
    """TDAA_beta/main_run_sstune_EvalVer.py:328-346 (the two debugging prints dropped)."""

    def __init__(self, flat=36480):
        super(Discriminator, self).__init__()
        self.cnn = nn.Conv2d(1, 64, (3, 3), stride=(2, 2), )
        self.cnn1 = nn.Conv2d(64, 64, (3, 3), stride=(2, 2), )
        self.cnn2 = nn.Conv2d(64, 64, (3, 3), stride=(2, 2), )
        self.final = nn.Linear(flat, 1)

    def forward(self, spec):
        bs, topk, len, fre = spec.size()
        spec = spec.view(bs * topk, 1, len, fre)
        spec = F.relu(self.cnn(spec))
        spec = F.relu(self.cnn1(spec))
        spec = F.relu(self.cnn2(spec))
        spec = spec.view(bs * topk, -1)
        score = torch.sigmoid(self.final(spec))
        return score


def gan_loss_terms_ref(score_true, score_false):
    """:643-652 and :670-671 with loss_dis_class = torch.nn.MSELoss() (:558)."""
    loss_dis_class = torch.nn.MSELoss()
    n = score_true.size()[0]
    acc_true = float(torch.sum(score_true > 0.5)) / float(n)
    acc_false = float(torch.sum(score_false < 0.5)) / float(n)
    loss_dis_true = loss_dis_class(score_true, torch.ones(n, 1))
    loss_dis_false = loss_dis_class(score_false, torch.zeros(n, 1))
    return {'loss_dis_true': loss_dis_true, 'loss_dis_false': loss_dis_false, 'loss_dis': loss_dis_true + loss_dis_false,
            'loss_gen': loss_dis_class(score_false, torch.ones(n, 1)),
            'acc_true': acc_true, 'acc_false': acc_false, 'acc_dis': (acc_false + acc_true) / 2}


class ATTENTION(nn.Module):
    def __init__(self, config, hidden_size, mode='dot'):
        super(ATTENTION, self).__init__()
        self.config = config
        self.hidden_size = hidden_size
        self.align_hidden_size = hidden_size
        self.mode = mode
        self.Linear_1 = nn.Linear(self.hidden_size, self.align_hidden_size, bias=False)
        self.Linear_2 = nn.Linear(hidden_size, self.align_hidden_size, bias=False)
        self.Linear_3 = nn.Linear(self.align_hidden_size, 1, bias=False)

    def forward(self, mix_hidden, query):
        BATCH_SIZE = mix_hidden.size()[0]
        if self.mode == 'dot':
            if not self.config.is_ComlexMask:
                mix_shape = mix_hidden.size()
                mix_hidden = mix_hidden.view(BATCH_SIZE, -1, self.hidden_size)
                query = query.view(-1, self.hidden_size, 1)
                dot = torch.baddbmm(torch.zeros(1, 1, dtype=mix_hidden.dtype), mix_hidden, query)
                energy = dot.view(BATCH_SIZE, mix_shape[1], mix_shape[2])
                return torch.sigmoid(energy)
            else:
                E = self.config.EMBEDDING_SIZE
                query_1 = query[:, :E].contiguous()
                query_2 = query[:, E:].contiguous()
                mix_shape = mix_hidden.size()
                mix_hidden = mix_hidden.view(BATCH_SIZE, -1, self.hidden_size)
                masks = []
                for q in (query_1, query_2):
                    q = q.view(-1, self.hidden_size, 1)
                    dot = torch.baddbmm(torch.zeros(1, 1, dtype=mix_hidden.dtype), mix_hidden, q)
                    energy = dot.view(BATCH_SIZE, mix_shape[1], mix_shape[2], 1)
                    masks.append(cRM_k * torch.tanh(energy))
                return torch.cat((masks[0], masks[1]), 3)
        elif self.mode == 'align':
            if self.config.is_ComlexMask:
                # broken in the reference (masks never appended, ...cRM_EvalVer.py:292-300)
                raise IndexError('cRM + align is not defined by the reference')
            mix_shape = mix_hidden.size()
            mix_hidden = mix_hidden.view(-1, self.hidden_size)
            mix_hidden = self.Linear_1(mix_hidden).view(BATCH_SIZE, -1, self.align_hidden_size)
            query = self.Linear_2(query).view(-1, 1, self.align_hidden_size)
            s = torch.tanh(mix_hidden + query)
            energy = self.Linear_3(s.view(-1, self.align_hidden_size)).view(
                BATCH_SIZE, mix_shape[1], mix_shape[2])
            return torch.sigmoid(energy)
        else:
            raise IndexError('NO this attention methods.')


class SPEECH_EMBEDDING(nn.Module):
    def __init__(self, config, num_labels, embedding_size, max_num_channel):
        super(SPEECH_EMBEDDING, self).__init__()
        self.num_all = num_labels
        self.emb_size = embedding_size
        self.max_num_out = max_num_channel
        if not config.is_ComlexMask:
            self.layer = nn.Embedding(num_labels, embedding_size)
        else:
            self.layer = nn.Embedding(num_labels, 2 * embedding_size)

    def forward(self, input, mask_idx):
        aim_matrix = torch.from_numpy(np.array(mask_idx, dtype=np.int64))
        return self.layer(aim_matrix)


class SPEECH_EMBEDDING_multihot(nn.Module):
    """Old form: every one of num_labels channels, zeroed where inactive
    (Torch_multi/main_run_multi_selfSS.py:308-328)."""

    def __init__(self, num_labels, embedding_size, max_num_channel):
        super(SPEECH_EMBEDDING_multihot, self).__init__()
        self.num_all = num_labels
        self.emb_size = embedding_size
        self.layer = nn.Embedding(num_labels, embedding_size)

    def forward(self, input):
        size = input.size()
        inp = input.long()
        order = torch.arange(0, self.num_all).view(1, self.num_all).repeat(size[0], 1)
        all_ = self.layer(order * inp)
        return all_ * input.view(size[0], size[1], 1).expand(size[0], size[1], self.emb_size).to(all_.dtype)


class ADDJUST(nn.Module):
    def __init__(self, config, hidden_units, embedding_size):
        super(ADDJUST, self).__init__()
        self.hidden_units = hidden_units
        if not config.is_ComlexMask:
            self.emb_size = embedding_size
        else:
            self.emb_size = 2 * embedding_size
        self.layer = nn.Linear(hidden_units + self.emb_size, self.emb_size, bias=False)

    def forward(self, input_hidden, prob_emb):
        B = input_hidden.size(0)
        top_k_num = prob_emb.size()[1]
        x = torch.mean(input_hidden, 1).view(B, 1, self.hidden_units).expand(B, top_k_num, self.hidden_units)
        can = torch.cat([x, prob_emb], dim=2)
        return self.layer(can)


def top_k_mask(batch_pro, alpha, top_k):
    size = batch_pro.size()
    final = torch.zeros(size)
    sort_result, sort_index = torch.sort(batch_pro, 1, True)
    sort_index = sort_index[:, :top_k]
    sort_result = torch.sum(sort_result > alpha, 1)
    for line_idx in range(size[0]):
        line_top_k = sort_index[line_idx][:int(sort_result[line_idx])]
        for i in line_top_k.numpy():
            final[line_idx, i] = 1
    return final


def recursive_extract_ref(config, classifier, mix_layer, emb_layer, att_layer, adj_layer, mix_feas, steps=3, alpha=-0.5):
    """Recursive extract-and-subtract inference, batched restatement of
    TDAA_beta/main_run_sstune_RecuVer.py:32-79 (`model_step_output`, test_mode: top_k = 1, alpha = -0.5) and the
    loop at :480-494: per step the classifier names the most probable remaining speaker, the attention path
    predicts that speaker's magnitude spectrogram from the CURRENT features, and the prediction is subtracted
    from the features before the next step.  mix_feas [B,T,F] -> (predict [B,steps,T,F], speakers [B,steps])."""
    now = mix_feas.clone()
    preds, spk = [], []
    for _ in range(steps):
        prob = classifier(now)
        mask = top_k_mask(prob, alpha, 1)
        idx = np.stack([np.where(line == 1)[0] for line in mask.numpy()])          # [B,1]
        out = forward_ref(config, mix_layer, emb_layer, att_layer, adj_layer, now, idx)
        step_pred = out['predict'][:, 0]
        preds.append(step_pred)
        spk.append(idx[:, 0])
        now = now - step_pred
    return torch.stack(preds, 1), np.stack(spk, 1)


# --------------------------------------------------------------------------------------
def forward_ref(config, mix_layer, emb_layer, att_layer, adj_layer, mix_feas, spk_idx, mix_mag=None):
    """The eval/train forward glue: features + speaker ids -> masks and predicted spectra.

    Follows TDAA_beta/main_run_sstune_EvalVer.py:420-470 (real masks) and
    TDAA_beta/main_run_sstune_cRM_EvalVer.py:498-553 (cRM), including the
    `expand(...).contiguous()` S-fold copy of the embedding tensor.
    mix_feas [B,T,F] float tensor, spk_idx int [B,S], mix_mag [B,T,F,2] (cRM only).
    Returns dict(masks, predict (real) | predict_real/predict_fake (cRM), hidden, query).
    """
    B, T, F = mix_feas.shape
    E = config.EMBEDDING_SIZE
    mix_speech_hidden, mix_tmp_hidden = mix_layer(mix_feas)
    embs = emb_layer(None, spk_idx)
    if adj_layer is not None:
        embs = adj_layer(mix_tmp_hidden, embs) + embs
    S = embs.size(1)
    h5 = mix_speech_hidden.view(B, 1, T, F, E).expand(B, S, T, F, E).contiguous().view(-1, T, F, E)
    out = {'hidden': mix_tmp_hidden, 'query': embs}
    if not config.is_ComlexMask:
        att = att_layer(h5, embs.view(-1, E)).view(B, S, T, F)
        out['masks'] = att
        out['predict'] = att * mix_feas.view(B, 1, T, F).expand(B, S, T, F)
    else:
        att = att_layer(h5, embs.view(-1, 2 * E)).view(B, S, T, F, 2)
        att = -1 / cRM_C * torch.log((cRM_k - att) / (cRM_k + att))
        out['masks'] = att
        x = mix_mag.view(B, 1, T, F, 2).expand(B, S, T, F, 2)
        mr, mi = att[..., 0], att[..., 1]
        xr, xi = x[..., 0], x[..., 1]
        out['predict_real'] = mr * xr - mi * xi
        out['predict_fake'] = mr * xi + mi * xr
    return out


def loss_ref(config, fwd, y_multi_map):
    """MSE losses: real TDAA_beta/main_run_sstune_EvalVer.py:487-497 ; cRM ...cRM_EvalVer.py:566-568."""
    mse = nn.MSELoss()
    if not config.is_ComlexMask:
        l1 = mse(fwd['predict'], y_multi_map)
        s = torch.sum(fwd['masks'], 1)
        l2 = mse(s, torch.ones_like(s))
        return l1 + 0.5 * l2, l1, l2
    lr = mse(fwd['predict_real'], y_multi_map[..., 0])
    li = mse(fwd['predict_fake'], y_multi_map[..., 1])
    return li + lr, lr, li


def pit_mse_ref(pred, target):
    """Brute-force permutation-invariant MSE (north-star extension; no reference counterpart,
    SURVEY F5).  pred/target [B,S,...] -> (mean over B of min-perm per-utterance MSE, perms [B,S])."""
    import itertools
    B, S = pred.shape[:2]
    p = pred.reshape(B, S, -1).double()
    t = target.reshape(B, S, -1).double()
    pair = ((p[:, :, None, :] - t[:, None, :, :]) ** 2).mean(-1)       # [B, S_pred, S_tgt]
    best = torch.full((B,), float('inf'), dtype=torch.float64)
    best_perm = torch.zeros(B, S, dtype=torch.long)
    for perm in itertools.permutations(range(S)):
        c = sum(pair[:, s, perm[s]] for s in range(S)) / S
        upd = c < best
        best = torch.where(upd, c, best)
        best_perm[upd] = torch.tensor(perm)
    return best.mean(), best_perm


def reconstruct_ref(fwd, mix_phase, hop, complex_mask=False):
    """bss_eval / bss_eval_cRM reconstruction (wav writing dropped):
    TDAA_beta/main_run_sstune_EvalVer.py:55-64 ; ...cRM_EvalVer.py:96-98.
    mix_phase complex [B,T,F] numpy.  Returns float32 [B,S,hop*(T-1)]."""
    from .stft_ref import istft_ref
    if not complex_mask:
        pred = fwd['predict'].detach().numpy()
        B, S = pred.shape[:2]
        outs = []
        for b in range(B):
            phase = np.angle(mix_phase[b])
            outs.append([istft_ref(np.transpose(pred[b, s] * np.exp(1j * phase)), hop) for s in range(S)])
        return np.array(outs, dtype=np.float32)
    pr = fwd['predict_real'].detach().numpy()
    pi = fwd['predict_fake'].detach().numpy()
    B, S = pr.shape[:2]
    return np.array([[istft_ref(np.transpose(pr[b, s] + 1j * pi[b, s]), hop) for s in range(S)]
                     for b in range(B)], dtype=np.float32)

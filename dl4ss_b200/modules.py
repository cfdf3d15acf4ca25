"""The reference's separation modules, same names / constructor arguments / forward layouts /
state-dict keys, running on the dl4ss_b200 CUDA kernels.

  MIX_SPEECH        TDAA_beta/main_run_sstune_EvalVer.py:277-303 (LSTM) ;
                    TDAA_beta/main_run_sstune_cRM_EvalVer.py:340-365 (GRU) ;
                    Torch_multi/test_multi_labels_speech.py:212-233 (LSTM 2x300)
  ATTENTION         TDAA_beta/main_run_sstune_EvalVer.py:199-242 ; cRM ...cRM_EvalVer.py:247-271
  SPEECH_EMBEDDING  TDAA_beta/main_run_sstune_EvalVer.py:348-361 ; cRM ...cRM_EvalVer.py:390-406
  ADDJUST           TDAA_beta/main_run_sstune_EvalVer.py:363-377 ; cRM ...cRM_EvalVer.py:408-426
  top_k_mask        TDAA_beta/main_run_sstune_EvalVer.py:390-405

Parameters live in ordinary torch containers (`self.layer = nn.LSTM(...)`, `self.Linear`, ...) so
`state_dict()` / `load_state_dict()` are key-compatible with the reference's checkpoints
(`layer.weight_ih_l0`, ..., `Linear.weight`, `Linear_1.weight`, ...); the containers' own forward
is never called -- every forward goes through the C ABI (dl4ss_b200._lib) and raises if the CUDA
library is missing.  Differences from the reference that do not change results: B is taken from
the input (the reference reads config.BATCH_SIZE inside forward), and MIX_SPEECH may hand
ATTENTION a `DeferredEmbedding` instead of the 8 MB/utterance [B,T,F,E] tensor (see below).
"""
import os
import numpy as np
import torch
from torch import nn

from . import _lib
from . import config


# ----------------------------------------------------------------------------- low-level ops
def linear_fwd(x2d, weight, bias=None, act='none', out=None):
    """act(x2d[M,K] @ weight[N,K]^T + bias) on the fp32 CUDA-core GEMM."""
    lib = _lib.load()
    M, K = x2d.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x2d.device, dtype=torch.float32)
    a = {'none': _lib.ACT_NONE, 'tanh': _lib.ACT_TANH, 'sigmoid': _lib.ACT_SIGMOID}[act]
    rc = lib.dl4ss_linear_fwd(_lib.ptr(x2d, name='x'), x2d.stride(0), _lib.ptr(weight, name='weight'),
                              weight.stride(0), _lib.ptr(bias, name='bias'), _lib.ptr(out), out.stride(0),
                              M, N, K, a, _lib.stream())
    _lib.check(rc, 'dl4ss_linear_fwd')
    return out


def split_bf16(x2d):
    """fp32 [R,K] -> bf16 hi/lo planes [2,R,Kp] (Kp = K rounded up to 64) for the tcgen05 GEMMs."""
    lib = _lib.load()
    R, K = x2d.shape
    Kp = (K + 63) // 64 * 64
    planes = torch.empty(2, R, Kp, device=x2d.device, dtype=torch.bfloat16)
    rc = lib.dl4ss_split_bf16(_lib.ptr(x2d, name='x'), x2d.stride(0), R, K, _lib.ptr(planes, torch.bfloat16),
                              _lib.stream())
    _lib.check(rc, 'dl4ss_split_bf16')
    return planes


def split_bf16_t(x2d):
    """fp32 [R,C] -> bf16 hi/lo planes of the TRANSPOSE, [2,C,Rp] (Rp = R rounded up to 64, zero padded)."""
    lib = _lib.load()
    R, C = x2d.shape
    Rp = (R + 63) // 64 * 64
    planes = torch.empty(2, C, Rp, device=x2d.device, dtype=torch.bfloat16)
    if x2d.stride(1) != 1 or x2d.dtype != torch.float32:
        x2d = x2d.contiguous().float()
    rc = lib.dl4ss_split_bf16_t(ctypes_ptr(x2d), x2d.stride(0), R, C,
                                _lib.ptr(planes, torch.bfloat16), _lib.stream())
    _lib.check(rc, 'dl4ss_split_bf16_t')
    return planes


def ctypes_ptr(t):
    """Device pointer of a row-strided (inner stride 1) CUDA fp32 tensor."""
    import ctypes
    if not t.is_cuda or t.dtype != torch.float32 or t.stride(-1) != 1:
        raise RuntimeError('dl4ss_b200: need a CUDA float32 tensor with unit inner stride')
    return ctypes.c_void_p(t.data_ptr())


def matmul_tn(a2d, b2d):
    """a2d^T @ b2d  ([R,Ca],[R,Cb] -> [Ca,Cb]) on tcgen05 (bf16x3): both operands transposed-split, contraction over R."""
    R = a2d.shape[0]
    return linear_tc(split_bf16_t(a2d), split_bf16_t(b2d), None, a2d.shape[1], b2d.shape[1], R, split_k=True)


def linear_tc_tn(a_planes, col0_a, M, shift_a, b_planes, col0_b, N, shift_b, B, T, wa=None, wb=None):
    """sum over (b,t) of A[b,t+shift_a,col0_a+m] * B[b,t+shift_b,col0_b+n] -> [M,N] fp32 from ROW-major bf16 hi/lo planes
    [2, B*T, ld] (MN-major UMMA operands: the weight gradients without a transposing split; out-of-range frames read zeros)."""
    lib = _lib.load()
    # a TMA box starts on a 16-byte boundary: windows that begin at an odd multiple of elements are widened to the left and
    # the surplus rows / columns of the result dropped
    pa, pb = col0_a % 8, col0_b % 8
    out = torch.empty(M + pa, N + pb, device=a_planes.device, dtype=torch.float32)
    lda, ldb = a_planes.shape[-1], b_planes.shape[-1]
    rc = lib.dl4ss_linear_tc_tn_splitk_fwd(_lib.ptr(a_planes, torch.bfloat16), lda, lda if wa is None else wa, col0_a - pa, shift_a,
                                           _lib.ptr(b_planes, torch.bfloat16), ldb, ldb if wb is None else wb, col0_b - pb, shift_b,
                                           _lib.ptr(out), out.stride(0), M + pa, N + pb, B, T, _lib.stream())
    _lib.check(rc, 'dl4ss_linear_tc_tn_splitk_fwd')
    return out[pa:, pb:] if (pa or pb) else out


def matmul_nn(a2d, w):
    """a2d @ w  ([M,K],[K,N] -> [M,N]) on tcgen05 (bf16x3): w is transposed once into the kernel's [N,K] operand form."""
    a2d = a2d if a2d.is_contiguous() else a2d.contiguous()
    return linear_tc(split_bf16(a2d), split_bf16(w.t().contiguous()), None, a2d.shape[0], w.shape[1], a2d.shape[1])


def weight_planes(w):
    """bf16 planes of a weight matrix, cached ON the tensor object and re-split only when the
    parameter is modified in place (optimizer step, load_state_dict) or moved."""
    ent = getattr(w, '_dl4ss_planes', None)
    if ent is None or ent[0] != w._version or ent[1] != w.data_ptr():
        ent = (w._version, w.data_ptr(), split_bf16(w.detach().contiguous()))
        w._dl4ss_planes = ent
    return ent[2]


def weight_t_planes(w):
    """bf16 planes of w^T ([K,N] weight used as the [N,K] operand of `a @ w`), cached like `weight_planes`."""
    ent = getattr(w, '_dl4ss_planes_t', None)
    if ent is None or ent[0] != w._version or ent[1] != w.data_ptr():
        ent = (w._version, w.data_ptr(), split_bf16(w.detach().t().contiguous()))
        w._dl4ss_planes_t = ent
    return ent[2]


def linear_tc(a_planes, w_planes, bias, M, N, K, out=None, act='none', split_k=False):
    """act(a[M,K] @ w[N,K]^T + bias) from pre-split planes on tcgen05 (bf16x3, fp32 accumulate).
    split_k: let the kernel cut K into splits that fill the SMs when the output is only a few tiles (act 'none')."""
    lib = _lib.load()
    if out is None:
        out = torch.empty(M, N, device=a_planes.device, dtype=torch.float32)
    if split_k and act == 'none':
        rc = lib.dl4ss_linear_tc_splitk_fwd(_lib.ptr(a_planes, torch.bfloat16), _lib.ptr(w_planes, torch.bfloat16),
                                            _lib.ptr(bias, name='bias'), _lib.ptr(out), out.stride(0), M, N, K,
                                            _lib.stream())
        _lib.check(rc, 'dl4ss_linear_tc_splitk_fwd')
        return out
    a = {'none': _lib.ACT_NONE, 'tanh': _lib.ACT_TANH, 'sigmoid': _lib.ACT_SIGMOID}[act]
    rc = lib.dl4ss_linear_tc_fwd(_lib.ptr(a_planes, torch.bfloat16), _lib.ptr(w_planes, torch.bfloat16),
                                 _lib.ptr(bias, name='bias'), _lib.ptr(out), out.stride(0), M, N, K, a, _lib.stream())
    _lib.check(rc, 'dl4ss_linear_tc_fwd')
    return out


def use_tensor_cores():
    return config.GEMM_PRECISION == 'bf16x3'


class _PackedRNN(object):
    """Per-layer operands of the recurrent kernel, rebuilt only when a parameter changes."""

    def __init__(self, rnn):
        self.rnn = rnn
        self.key = None
        self.layers = None

    def get(self):
        rnn = self.rnn
        key = tuple((p.data_ptr(), p._version) for p in rnn.parameters())
        if key == self.key:
            return self.layers
        H = rnn.hidden_size
        gru = isinstance(rnn, nn.GRU)
        layers = []
        with torch.no_grad():
            for l in range(rnn.num_layers):
                wih, bias, whh, bhn = [], [], [], []
                for suf in ('', '_reverse'):
                    w_ih = getattr(rnn, 'weight_ih_l%d%s' % (l, suf))
                    w_hh = getattr(rnn, 'weight_hh_l%d%s' % (l, suf))
                    b_ih = getattr(rnn, 'bias_ih_l%d%s' % (l, suf))
                    b_hh = getattr(rnn, 'bias_hh_l%d%s' % (l, suf))
                    wih.append(w_ih)
                    whh.append(w_hh)
                    if gru:      # n-gate: b_hn stays inside r*(W_hn h + b_hn)
                        b = b_ih.clone()
                        b[:2 * H] += b_hh[:2 * H]
                        bias.append(b)
                        bhn.append(b_hh[2 * H:])
                    else:
                        bias.append(b_ih + b_hh)
                layers.append({
                    'wih': torch.cat(wih, 0).contiguous().float(),
                    'bias': torch.cat(bias, 0).contiguous().float(),
                    'whh': torch.stack(whh, 0).contiguous().float(),
                    'bhn': torch.stack(bhn, 0).contiguous().float() if gru else None,
                })
        self.key, self.layers = key, layers
        return layers


def use_tc_recurrence(H, cell):
    """tcgen05 recurrent kernel: bf16x3 mode, supported H (multiple of 4, <= 320), not disabled."""
    return (use_tensor_cores() and config.RNN_TENSOR_CORES
            and bool(_lib.load().dl4ss_rnn_tc_supported(H, cell)))


def use_mma_recurrence(H, cell):
    """Warp-level tensor-core recurrent kernel (bf16x3 mma.sync) for hidden sizes the tcgen05 form cannot hold (H = 600)."""
    return (use_tensor_cores() and config.RNN_TENSOR_CORES and not use_tc_recurrence(H, cell)
            and bool(_lib.load().dl4ss_rnn_mma_supported(H, cell)))


def recurrent_workspace(B, T, H, cell, tc_rec, device):
    """tc_rec: True (tcgen05 kernel), 'mma' (warp-level tensor-core kernel) or False (fp32 kernel)."""
    lib = _lib.load()
    if tc_rec == 'mma':
        n = int(lib.dl4ss_rnn_mma_workspace_bytes(B, T, H, cell))
    else:
        n = int(lib.dl4ss_rnn_tc_workspace_bytes(B, T, H, cell) if tc_rec else lib.dl4ss_rnn_workspace_bytes(B, T, H, cell))
    return torch.empty(n, device=device, dtype=torch.uint8)


def whh_planes(lw, cell, H):
    """Packed bf16 hi/lo planes of W_hh for the tcgen05 recurrent kernel (built once per weight version:
    `lw` is rebuilt by _PackedRNN whenever a parameter changes)."""
    pl = lw.get('whh_tc')
    if pl is None:
        lib = _lib.load()
        pl = torch.empty(int(lib.dl4ss_rnn_tc_whh_bytes(H)) // 2, device=lw['whh'].device, dtype=torch.bfloat16)
        rc = lib.dl4ss_rnn_tc_pack_whh(cell, _lib.ptr(lw['whh']), H, _lib.ptr(pl, torch.bfloat16), _lib.stream())
        _lib.check(rc, 'dl4ss_rnn_tc_pack_whh')
        lw['whh_tc'] = pl
    return pl


def recurrent_layer(lw, cell, xproj, B, T, H, ws, tc_rec, gates=None, cells=None, y=None, y_planes=None, hmean=None,
                    want_y=True):
    """K3: one bidirectional layer over the hoisted input projection xproj [B*T, 2*G*H] -> y [B,T,2H]
    (want_y=False with y_planes on the tcgen05 kernel: planes only, returns None)."""
    lib = _lib.load()
    if y is None and (want_y or y_planes is None or tc_rec is not True):
        y = torch.empty(B, T, 2 * H, device=xproj.device, dtype=torch.float32)
    if tc_rec == 'mma':
        rc = lib.dl4ss_rnn_layer_mma_fwd(cell, _lib.ptr(xproj), _lib.ptr(lw['whh']), _lib.ptr(lw['bhn']),
                                         _lib.ptr(y), B, T, H, _lib.ptr(gates), _lib.ptr(cells),
                                         _lib.ptr(ws, torch.uint8), ws.numel(), _lib.stream())
        _lib.check(rc, 'dl4ss_rnn_layer_mma_fwd')
    elif tc_rec:
        rc = lib.dl4ss_rnn_layer_tc_fwd(cell, _lib.ptr(xproj), _lib.ptr(whh_planes(lw, cell, H), torch.bfloat16),
                                        _lib.ptr(lw['bhn']), _lib.ptr(y), B, T, H, _lib.ptr(gates), _lib.ptr(cells),
                                        _lib.ptr(y_planes, torch.bfloat16), _lib.ptr(hmean),
                                        _lib.ptr(ws, torch.uint8), ws.numel(), _lib.stream())
        _lib.check(rc, 'dl4ss_rnn_layer_tc_fwd')
    else:
        rc = lib.dl4ss_rnn_layer_fwd(cell, _lib.ptr(xproj), _lib.ptr(lw['whh']), _lib.ptr(lw['bhn']),
                                     _lib.ptr(y), B, T, H, _lib.ptr(gates), _lib.ptr(cells),
                                     _lib.ptr(ws, torch.uint8), ws.numel(), _lib.stream())
        _lib.check(rc, 'dl4ss_rnn_layer_fwd')
    return y


_RNN_YX = os.environ.get('DL4SS_RNN_YX', '1') != '0'      # csrc/rnn_tc.cu: h exchanged through the output planes (the library's default)


class HiddenStub(object):
    """Shape / device of an encoder output that was produced as bf16 planes only (`rnn_forward(..., need_y=False)`)."""

    def __init__(self, shape, device):
        self.shape, self.device = torch.Size(shape), device

    def view(self, *a):
        raise RuntimeError('the encoder output exists as bf16 planes only (extras["planes"]); call encode() without planes_only')


def rnn_forward(packed, x, save=None, buffers=None, extras=None, need_y=True):
    """Bidirectional multi-layer LSTM/GRU forward, batch_first, zero initial state.
    x [B,T,in] -> y [B,T,2H].  `save` (list) receives per-layer tensors for backward; `buffers` (list of dicts
    with 'y', 'gates', 'cells' per layer) makes the layers write into caller-owned static tensors; `extras` (dict)
    receives what the last layer's kernel produced on the side: 'planes' (its output as bf16 hi/lo planes, ready
    for the next tensor-core projection) and 'hmean' (its mean over T), or None when the fp32 kernels ran."""
    lib = _lib.load()
    rnn = packed.rnn
    gru = isinstance(rnn, nn.GRU)
    cell = _lib.CELL_GRU if gru else _lib.CELL_LSTM
    G = 3 if gru else 4
    H = rnn.hidden_size
    B, T, _ = x.shape
    dev = x.device
    tc_rec = use_tc_recurrence(H, cell)
    if not tc_rec and use_mma_recurrence(H, cell):
        tc_rec = 'mma'
    ws = recurrent_workspace(B, T, H, cell, tc_rec, dev)
    inp = x.contiguous()
    xproj = torch.empty(B * T, 2 * G * H, device=dev, dtype=torch.float32)
    layers = packed.get()
    fuse = tc_rec is True and use_tensor_cores()        # K3 emits its output pre-split for the next projection (+ the T-mean)
    planes = None                               # bf16 hi/lo planes of `inp`, when the producer made them
    hmean = None
    Kpy = (2 * H + 63) // 64 * 64
    for li, lw in enumerate(layers):
        a_pl = None
        if use_tensor_cores():
            a_pl = planes if planes is not None else split_bf16(inp.view(B * T, -1))
            linear_tc(a_pl, weight_planes(lw['wih']), lw['bias'], B * T, 2 * G * H, inp.shape[-1], out=xproj)
        else:
            linear_fwd(inp.view(B * T, -1), lw['wih'], lw['bias'], 'none', out=xproj)
        gates = cells = y_out = None
        if buffers is not None:
            gates, cells, y_out = buffers[li]['gates'], buffers[li]['cells'], buffers[li]['y']
        elif save is not None:
            gates = torch.empty(B, T, 2, G * H, device=dev, dtype=torch.float32)
            cells = torch.empty(B, T, 2, H, device=dev, dtype=torch.float32)
        planes = None
        if fuse:
            planes = torch.empty(2, B * T, Kpy, device=dev, dtype=torch.bfloat16)
            if Kpy > 2 * H and not _RNN_YX:
                planes[:, :, 2 * H:].zero_()    # the kernel writes columns [0, 2H) (and, exchanging h through the planes, zeroes the padding itself)
            if li == len(layers) - 1:
                hmean = torch.empty(B, 2 * H, device=dev, dtype=torch.float32)
        skip_y = fuse and not need_y and save is None and buffers is None
        y = recurrent_layer(lw, cell, xproj, B, T, H, ws, tc_rec, gates, cells, y_out, planes,
                            hmean if li == len(layers) - 1 else None, want_y=not skip_y)
        if y is None:
            y = HiddenStub((B, T, 2 * H), dev)
        if save is not None:
            save.append({'x': inp, 'y': y, 'gates': gates, 'cells': cells,
                         'x_planes': a_pl if use_tensor_cores() else None, 'y_planes': planes})
        inp = y
    if extras is not None:
        extras['planes'] = planes
        extras['hmean'] = hmean
    return inp


def emb_attn_mask(h, weight, bias, query, F, E, complex_mask=False, decompress=True, h_planes=None):
    """Fused Linear+tanh -> dot attention -> masks (K4).
    h [B,T,K], weight [F*E,K], bias [F*E], query [B,S,E|2E] -> [B,S,T,F] (or [B,S,T,F,2])."""
    lib = _lib.load()
    B, T, K = h.shape
    S = query.shape[1]
    dev = h.device
    mode = _lib.ATT_DOT_CRM if complex_mask else _lib.ATT_DOT
    out = torch.empty((B, S, T, F, 2) if complex_mask else (B, S, T, F), device=dev, dtype=torch.float32)
    query = query.contiguous()
    if use_tensor_cores() and E == 50 and S <= 4:
        # keep both plane tensors referenced until the launch is queued: a temporary freed between two argument
        # expressions can be handed to the next allocation (e.g. the first-use weight split) and overwritten
        w_pl = weight_planes(weight)
        h_pl = h_planes if h_planes is not None else split_bf16(h.view(B * T, K))
        rc = lib.dl4ss_emb_attn_mask_tc_fwd(_lib.ptr(h_pl, torch.bfloat16),
                                            _lib.ptr(w_pl, torch.bfloat16),
                                            _lib.ptr(bias, name='bias'), _lib.ptr(query, name='query'),
                                            B, T, F, E, K, S, mode, float(config.cRM_k),
                                            float(config.cRM_C if decompress else 0.0), _lib.ptr(out), _lib.stream())
        _lib.check(rc, 'dl4ss_emb_attn_mask_tc_fwd')
        return out
    ws_bytes = int(lib.dl4ss_emb_attn_mask_workspace_bytes(B, T, F, E))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    rc = lib.dl4ss_emb_attn_mask_fwd(_lib.ptr(h, name='h'), _lib.ptr(weight, name='weight'),
                                     _lib.ptr(bias, name='bias'), _lib.ptr(query, name='query'),
                                     B, T, F, E, K, S, mode, float(config.cRM_k),
                                     float(config.cRM_C if decompress else 0.0), _lib.ptr(out),
                                     _lib.ptr(ws, torch.uint8), ws_bytes, _lib.stream())
    _lib.check(rc, 'dl4ss_emb_attn_mask_fwd')
    return out


def crm_decompress(att):
    """-1/C * log((K-m)/(K+m))  (TDAA_beta/main_run_sstune_cRM_EvalVer.py:512); plain torch glue,
    only used when a caller keeps the reference's two-step form."""
    return -1 / config.cRM_C * torch.log((config.cRM_k - att) / (config.cRM_k + att))


class _TrainsThroughTrainStep(torch.autograd.Function):
    """Identity whose backward explains where the training path is.  The drop-in modules run hand-written forward
    kernels on detached weights; the matching backward kernels are driven by `dl4ss_b200.TrainStep`, not by autograd."""

    @staticmethod
    def forward(ctx, x, anchor):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        raise RuntimeError('dl4ss_b200: the drop-in modules have no autograd graph -- loss.backward() cannot reach their '
                           'parameters.  Run training steps through dl4ss_b200.TrainStep (forward + hand-written backward '
                           'kernels + optimizer step), see INTEGRATION.md.')


def _mark_no_autograd(out, param):
    """Outside torch.no_grad() with trainable parameters: hang the loud backward on the output (evaluation loops that
    never call backward are unaffected)."""
    if torch.is_grad_enabled() and param.requires_grad and torch.is_tensor(out):
        return _TrainsThroughTrainStep.apply(out, param)
    return out


# ----------------------------------------------------------------------------- deferred tensor
class DeferredEmbedding(object):
    """Stand-in for MIX_SPEECH's [B,T,F,E] output that is never written to HBM.

    It answers the shape calls the reference's glue makes on that tensor
    (`.view(B,1,T,F,E).expand(B,S,T,F,E).contiguous().view(-1,T,F,E)`,
    TDAA_beta/main_run_sstune_EvalVer.py:453-455) and is consumed by ATTENTION.forward, which then
    runs the fused Linear+tanh+attention kernel on the encoder output.  `.materialize()` returns
    the real tensor (costing the 8 MB/utterance the reference always pays)."""

    def __init__(self, hidden, weight, bias, F, E, shape=None):
        self.hidden, self.weight, self.bias = hidden, weight, bias
        self.F, self.E = F, E
        B, T, _ = hidden.shape
        self.base = (B, T, F, E)
        self._shape = tuple(shape) if shape is not None else self.base

    def _like(self, shape):
        shape = list(shape)
        n = int(np.prod(self.base))
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1]))
            total = int(np.prod(self._shape))
            shape[shape.index(-1)] = total // known
        if int(np.prod(shape)) % n != 0 or tuple(shape[-3:]) != self.base[1:]:
            raise RuntimeError('DeferredEmbedding only supports the [B(,S),T,F,E] reshapes of the '
                               'reference glue; call .materialize() for anything else')
        return DeferredEmbedding(self.hidden, self.weight, self.bias, self.F, self.E, shape)

    def view(self, *shape):
        return self._like(shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape)

    reshape = view

    def expand(self, *shape):
        return self._like(shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape)

    def contiguous(self):
        return self

    def size(self, dim=None):
        return torch.Size(self._shape) if dim is None else self._shape[dim]

    @property
    def shape(self):
        return torch.Size(self._shape)

    def dim(self):
        return len(self._shape)

    @property
    def copies(self):
        """S of the expand: how many speaker queries share each utterance's embedding."""
        return int(np.prod(self._shape)) // int(np.prod(self.base))

    def materialize(self):
        B, T, F, E = self.base
        out = linear_fwd(self.hidden.view(B * T, -1), self.weight, self.bias, 'tanh').view(B, T, F, E)
        S = self.copies
        if S > 1:
            out = out.view(B, 1, T, F, E).expand(B, S, T, F, E).contiguous()
        return out.view(self._shape)


# ----------------------------------------------------------------------------- modules
class MIX_SPEECH(nn.Module):
    """Mixture encoder: bidirectional LSTM/GRU -> Linear(2H -> F*E) -> tanh -> [B,T,F,E].

    MIX_SPEECH(input_fre, mix_speech_len) as in the reference; extra keyword arguments select the
    variant the different reference scripts hard-code: cell 'lstm' | 'gru', num_layers (default
    config.NUM_LAYERS; the TDAA LSTM scripts hard-code 4), return_hidden (TDAA variants return
    `(out, rnn_output)`, Torch_multi returns `out`), fused (hand ATTENTION a DeferredEmbedding)."""

    def __init__(self, input_fre, mix_speech_len, cell='lstm', num_layers=None, return_hidden=True, fused=True):
        super(MIX_SPEECH, self).__init__()
        self.input_fre = input_fre
        self.mix_speech_len = mix_speech_len
        self.return_hidden = return_hidden
        self.fused = fused
        rnn = {'lstm': nn.LSTM, 'gru': nn.GRU}[cell]
        self.layer = rnn(input_size=input_fre, hidden_size=config.HIDDEN_UNITS,
                         num_layers=config.NUM_LAYERS if num_layers is None else num_layers,
                         batch_first=True, bidirectional=True)
        self.Linear = nn.Linear(2 * config.HIDDEN_UNITS, self.input_fre * config.EMBEDDING_SIZE)
        self._packed = _PackedRNN(self.layer)

    def encode(self, x, extras=None, planes_only=False):
        """x [B,T,F] -> hidden [B,T,2H].  planes_only (needs `extras`): on the fused tensor-core path the layers emit bf16 planes
        only (extras['planes'], extras['hmean']) and a `HiddenStub` with the shape is returned -- what `Separator` consumes."""
        return rnn_forward(self._packed, x, extras=extras, need_y=not (planes_only and extras is not None))

    def forward(self, x):
        B, T, F = x.shape
        xx = self.encode(x)
        E = self.Linear.out_features // self.input_fre
        if self.fused:
            out = DeferredEmbedding(xx, self.Linear.weight, self.Linear.bias, F, E)
        else:
            out = linear_fwd(xx.view(B * T, -1), self.Linear.weight.detach(), self.Linear.bias.detach(),
                             'tanh').view(B, T, F, E)
        xx = _mark_no_autograd(xx, self.Linear.weight)
        out = _mark_no_autograd(out, self.Linear.weight)
        return (out, xx) if self.return_hidden else out


class MIX_SPEECH_classifier(nn.Module):
    """MIX_SPEECH_classifier(input_fre, mix_speech_len, num_labels).forward(x[B,T,F]) -> speaker probabilities
    [B,num_labels]: BLSTM 3 x (2*HIDDEN_UNITS) -> mean over T -> Linear -> sigmoid
    (TDAA_beta/main_run_sstune_EvalVer.py:305-326).  The step before the separation path at
    inference: `top_k_mask` of its output picks the speakers to extract.  H = 600 is outside the TMEM-resident
    recurrent kernel's range (W_hh would need 600 TMEM columns per plane); its layers run on the batch-as-N tensor-core
    recurrent kernel (dl4ss_rnn_layer_mma_fwd).  Inference only: there is no training path for the classifier."""

    def __init__(self, input_fre, mix_speech_len, num_labels):
        super(MIX_SPEECH_classifier, self).__init__()
        self.input_fre = input_fre
        self.mix_speech_len = mix_speech_len
        self.layer = nn.LSTM(input_size=input_fre, hidden_size=2 * config.HIDDEN_UNITS, num_layers=3,
                             batch_first=True, bidirectional=True)
        self.Linear = nn.Linear(2 * 2 * config.HIDDEN_UNITS, num_labels)
        self._packed = _PackedRNN(self.layer)

    def forward(self, x):
        extras = {}
        y = rnn_forward(self._packed, x, extras=extras)
        m = extras.get('hmean')
        if m is None:
            m = y.mean(1)                      # [B, 4*HIDDEN_UNITS]
        out = linear_fwd(m.contiguous(), self.Linear.weight.detach(), self.Linear.bias.detach(), 'sigmoid')
        return _mark_no_autograd(out, self.Linear.weight)


class Discriminator(nn.Module):
    """Discriminator().forward(spec[bs,topk,len,fre]) -> score [bs*topk, 1]: real (clean-source) or predicted spectrogram
    (TDAA_beta/main_run_sstune_EvalVer.py:328-346): conv 3x3 stride 2 (1->64) + ReLU, twice more (64->64), flatten,
    Linear(36480 -> 1), sigmoid.  36480 = 64 x 38 x 15 is what a [313,129] spectrogram leaves; other sizes need a matching
    `final` layer (`flat_features(len, fre)`).  State-dict keys as the reference's: cnn / cnn1 / cnn2 / final.
    Forward only (dl4ss_conv3x3s2_relu_fwd + dl4ss_rowdot_sigmoid_fwd): the least-squares GAN terms the reference adds
    with it are `gan_loss_terms`; their gradients are not part of the C ABI (the backward raises, as for every module)."""

    def __init__(self, flat=36480):
        super(Discriminator, self).__init__()
        self.cnn = nn.Conv2d(1, 64, (3, 3), stride=(2, 2))
        self.cnn1 = nn.Conv2d(64, 64, (3, 3), stride=(2, 2))
        self.cnn2 = nn.Conv2d(64, 64, (3, 3), stride=(2, 2))
        self.final = nn.Linear(flat, 1)

    @staticmethod
    def flat_features(length, fre):
        h, w = length, fre
        for _ in range(3):
            h, w = (h - 3) // 2 + 1, (w - 3) // 2 + 1
        return 64 * h * w

    def forward(self, spec):
        lib = _lib.load()
        bs, topk, length, fre = spec.shape
        x = spec.contiguous().view(bs * topk, 1, length, fre)
        n = bs * topk
        for conv in (self.cnn, self.cnn1, self.cnn2):
            cin, ih, iw = x.shape[1:]
            oh, ow = (ih - 3) // 2 + 1, (iw - 3) // 2 + 1
            y = torch.empty(n, conv.out_channels, oh, ow, device=x.device, dtype=torch.float32)
            rc = lib.dl4ss_conv3x3s2_relu_fwd(_lib.ptr(x, name='spec'), _lib.ptr(conv.weight.detach(), name='weight'),
                                              _lib.ptr(conv.bias.detach()), _lib.ptr(y), n, cin, ih, iw, conv.out_channels,
                                              _lib.stream())
            _lib.check(rc, 'dl4ss_conv3x3s2_relu_fwd')
            x = y
        k = x[0].numel()
        if k != self.final.in_features:
            raise RuntimeError('size mismatch: the convolutions leave %d features, `final` expects %d' % (k, self.final.in_features))
        score = torch.empty(n, 1, device=x.device, dtype=torch.float32)
        rc = lib.dl4ss_rowdot_sigmoid_fwd(_lib.ptr(x), _lib.ptr(self.final.weight.detach(), name='final.weight'),
                                          _lib.ptr(self.final.bias.detach()), _lib.ptr(score), n, k, _lib.stream())
        _lib.check(rc, 'dl4ss_rowdot_sigmoid_fwd')
        return _mark_no_autograd(score, self.final.weight)


def gan_loss_terms(score_true, score_false):
    """The discriminator terms of the reference's training loop (TDAA_beta/main_run_sstune_EvalVer.py:643-652,670-671;
    `loss_dis_class` is MSELoss, :558): returns a dict of 0-d tensors
      loss_dis_true  = MSE(score_true, 1)     loss_dis_false = MSE(score_false, 0)     loss_dis = their sum (discriminator step)
      loss_gen       = MSE(score_false, 1)    (added to the separation loss in the generator step, :670-671)
      acc_true / acc_false / acc_dis : the accuracies it prints (:645-647)."""
    st, sf = score_true.detach().float(), score_false.detach().float()
    t = ((st - 1.0) ** 2).mean()
    f = (sf ** 2).mean()
    n = float(st.shape[0])
    acc_t = (st > 0.5).sum() / n
    acc_f = (sf < 0.5).sum() / n
    return {'loss_dis_true': t, 'loss_dis_false': f, 'loss_dis': t + f, 'loss_gen': ((sf - 1.0) ** 2).mean(),
            'acc_true': acc_t, 'acc_false': acc_f, 'acc_dis': (acc_t + acc_f) / 2}


class ATTENTION(nn.Module):
    """ATTENTION(hidden_size, mode='dot'|'align').forward(mix_hidden[N,T,F,E], query[N,E|2E]).

    Returns mask [N,T,F]; with config.is_ComlexMask, K*tanh(.) pairs [N,T,F,2] (decompression is
    the caller's next line in the reference, kept that way here).  Linear_1/2/3 exist in every
    mode, as in the reference, so checkpoints load."""

    def __init__(self, hidden_size, mode='dot'):
        super(ATTENTION, self).__init__()
        self.hidden_size = hidden_size
        self.align_hidden_size = hidden_size
        self.mode = mode
        self.Linear_1 = nn.Linear(self.hidden_size, self.align_hidden_size, bias=False)
        self.Linear_2 = nn.Linear(hidden_size, self.align_hidden_size, bias=False)
        self.Linear_3 = nn.Linear(self.align_hidden_size, 1, bias=False)

    def forward(self, mix_hidden, query):
        lib = _lib.load()
        cplx = bool(config.is_ComlexMask)
        E = self.hidden_size
        if self.mode == 'dot':
            if isinstance(mix_hidden, DeferredEmbedding):
                B, T, F, _ = mix_hidden.base
                S = mix_hidden.copies
                q = query.contiguous().view(B, S, -1)
                out = emb_attn_mask(mix_hidden.hidden, mix_hidden.weight, mix_hidden.bias, q, F, E,
                                    complex_mask=cplx, decompress=False)
                return _mark_no_autograd(out.view((B * S, T, F, 2) if cplx else (B * S, T, F)), mix_hidden.weight)
            N, T, F, _ = mix_hidden.shape
            mix_hidden = mix_hidden.contiguous()
            q = query.contiguous().view(N, 1, -1)
            out = torch.empty((N, T, F, 2) if cplx else (N, T, F), device=mix_hidden.device, dtype=torch.float32)
            rc = lib.dl4ss_attn_dot_fwd(_lib.ptr(mix_hidden, name='mix_hidden'), T * F * E, _lib.ptr(q, name='query'),
                                        N, 1, T * F, E, _lib.ATT_DOT_CRM if cplx else _lib.ATT_DOT,
                                        float(config.cRM_k), 0.0, _lib.ptr(out), _lib.stream())
            _lib.check(rc, 'dl4ss_attn_dot_fwd')
            return out
        elif self.mode == 'align':
            if cplx:
                # the reference's cRM+align branch never fills `masks` (cRM_EvalVer.py:292-300)
                raise IndexError('cRM + align attention is undefined in the reference')
            if isinstance(mix_hidden, DeferredEmbedding):
                mix_hidden = mix_hidden.materialize()
            N, T, F, _ = mix_hidden.shape
            a = linear_fwd(mix_hidden.contiguous().view(-1, E), self.Linear_1.weight.detach()).view(N, T * F, -1)
            qq = linear_fwd(query.contiguous().view(N, E), self.Linear_2.weight.detach()).view(N, 1, -1)
            s = torch.tanh(a + qq)
            en = linear_fwd(s.view(-1, self.align_hidden_size), self.Linear_3.weight.detach(), None, 'sigmoid')
            return en.view(N, T, F)
        else:
            raise IndexError('NO this attention methods.')


def _speaker_query(h, table, idx, wadj, residual):
    lib = _lib.load()
    dev = table.device
    if idx is not None:
        B, S = idx.shape
        EQ = table.shape[1]
        num = table.shape[0]
    else:
        B, S, EQ = table.shape
        num = 1
    q = torch.empty(B, S, EQ, device=dev, dtype=torch.float32)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    T = C = 1
    if h is not None:
        _, T, C = h.shape
    rc = lib.dl4ss_speaker_query_fwd(_lib.ptr(h, name='hidden'), B, T, C, _lib.ptr(table, name='table'), num, EQ,
                                     _lib.ptr(idx, torch.int64, 'idx'), S, _lib.ptr(wadj, name='adjust weight'),
                                     int(residual), _lib.ptr(q), None, _lib.ptr(err, torch.int32), _lib.stream())
    _lib.check(rc, 'dl4ss_speaker_query_fwd')
    return q, err


class SPEECH_EMBEDDING(nn.Module):
    """SPEECH_EMBEDDING(num_labels, embedding_size, max_num_channel).forward(input, mask_idx)
    -> [B,S,E] (2E wide when config.is_ComlexMask).  `input` is ignored, as in the reference."""

    def __init__(self, num_labels, embedding_size, max_num_channel):
        super(SPEECH_EMBEDDING, self).__init__()
        self.num_all = num_labels
        self.emb_size = embedding_size
        self.max_num_out = max_num_channel
        if not config.is_ComlexMask:
            self.layer = nn.Embedding(num_labels, embedding_size)
        else:
            self.layer = nn.Embedding(num_labels, 2 * embedding_size)

    def index_tensor(self, mask_idx):
        if torch.is_tensor(mask_idx):
            return mask_idx.to(device=self.layer.weight.device, dtype=torch.int64).contiguous()
        return torch.from_numpy(np.array(mask_idx, dtype=np.int64)).to(self.layer.weight.device)

    def forward(self, input, mask_idx=None):
        if mask_idx is None:
            return self.forward_multihot(input)
        idx = self.index_tensor(mask_idx)
        q, err = _speaker_query(None, self.layer.weight.detach(), idx, None, 1)
        if int(err.item()):
            raise IndexError('index out of range in self')     # nn.Embedding's error
        return q

    def forward_multihot(self, input):
        """The older one-argument form (Torch_multi/main_run_multi_selfSS.py:308-328): `input` [B,num_labels] is the
        0/1 `top_k_mask`; every one of the num_labels channels gets a vector -- row (channel index x input) of the
        table, i.e. the channel's own row where it is active and row 0 elsewhere -- multiplied by `input`, so
        inactive channels are exactly zero.  -> [B,num_labels,E]."""
        dev = self.layer.weight.device
        inp = input.to(device=dev, dtype=torch.float32)
        order = torch.arange(self.num_all, device=dev, dtype=torch.int64).unsqueeze(0)
        idx = (order * inp.to(torch.int64)).contiguous()
        q, _ = _speaker_query(None, self.layer.weight.detach(), idx, None, 1)
        return q * inp.unsqueeze(-1)


class ADDJUST(nn.Module):
    """ADDJUST(hidden_units, embedding_size).forward(input_hidden[B,T,2H], prob_emb[B,S,E])
    -> Linear([mean_T(input_hidden) ; prob_emb]) (no bias); the caller adds the residual."""

    def __init__(self, hidden_units, embedding_size):
        super(ADDJUST, self).__init__()
        self.hidden_units = hidden_units
        self.emb_size = embedding_size if not config.is_ComlexMask else 2 * embedding_size
        self.layer = nn.Linear(hidden_units + self.emb_size, self.emb_size, bias=False)

    def forward(self, input_hidden, prob_emb):
        q, _ = _speaker_query(input_hidden.contiguous(), prob_emb.contiguous(), None,
                              self.layer.weight.detach(), 0)
        return q


def top_k_mask(batch_pro, alpha, top_k):
    """Speaker selection (TDAA_beta/main_run_sstune_EvalVer.py:390-405): per row, the (at most) `top_k` most
    probable speakers whose probability exceeds `alpha`, as a 0/1 matrix.  Same result as the reference's
    python loops, evaluated with tensor ops on the device `batch_pro` lives on."""
    vals, index = torch.sort(batch_pro.detach(), 1, True)
    keep = (vals[:, :top_k] > alpha).to(torch.float32)
    final = torch.zeros(batch_pro.size(), device=batch_pro.device, dtype=torch.float32)
    final.scatter_(1, index[:, :top_k], keep)
    return final

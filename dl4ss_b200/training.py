"""Training step of the separation path on the GPU: forward with saved activations, loss, backward,
(sharded) gradient all-reduce, optimizer step.

Mirrors one iteration of the reference's training loop
(TDAA_beta/main_run_sstune_EvalVer.py:586-675 ; cRM: TDAA_beta/main_run_sstune_cRM_EvalVer.py:645-753):
    feas -> MIX_SPEECH -> SPEECH_EMBEDDING (+ADDJUST) -> ATTENTION -> masks -> mask*mix -> MSELoss
    (+0.5*MSE(sum_s mask, 1)) -> zero_grad -> backward -> Adam.step
without the discriminator terms (out of scope).  The forward runs the same CUDA kernels
as inference (tcgen05 projections, persistent recurrent kernel) with the gate activations saved; the
backward is written out by hand: fused CUDA kernels for the loss / attention / BPTT gate arithmetic
(csrc/train.cu), the tcgen05 bf16x3 projection kernel for the dense contractions dW = dY^T X (operands through
the transposing split dl4ss_split_bf16_t) and dX = dY W, and a plain library GEMM (torch.bmm -> cuBLAS fp32)
only for the tiny per-step dh_rec = dgates x W_hh of the BPTT chain.  Gradients land in the parameters'
`.grad`, so any torch optimizer (the reference uses Adam, lr 2e-4) steps them.

Utterances shard by batch across ranks: every rank runs the step on its shard with the
loss normalised by the GLOBAL element count, then one all-reduce (sum) of the flat fp32 gradient
bucket reproduces the global-batch MSELoss gradients on every rank.
"""
import ctypes

import torch
from torch import nn

from . import _lib
from . import config
from . import modules as M


def shard_range(n, rank, world):
    """Contiguous utterance shard [lo, hi) of rank `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_batch_size(local_b, device, group=None):
    """Utterances of the whole step over all ranks (sum of the shard sizes; the local size without a process group):
    the MSELoss means -- and so the summed gradients -- must be normalised by THIS, not by the shard size."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(local_b)
    n = torch.tensor([int(local_b)], device=device, dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return int(n.item())


class GradBucket(object):
    """ONE persistent flat fp32 gradient buffer; every parameter's `.grad` is a VIEW into it, so gradients are written
    in place by the backward kernels and reduced in place -- no flatten / unflatten copies, no per-step allocation.

    The buffer is cut into SEGMENTS in the order the backward pass finishes them (the head: Linear + embedding +
    ADDJUST, then the recurrent layers top-down).  `reduce_async(k)` launches the all-reduce (sum) of segment k as soon
    as its gradients are written: the collective runs on the process group's own stream (NCCL over NVLink on the GPU
    box, gloo in the CPU tests) underneath the BPTT chain and weight-gradient GEMMs of the layers below; `wait()` makes
    the current stream wait for all of them before the optimizer step (bucket per layer to overlap with the
    BPTT of lower layers)."""

    def __init__(self, segments):
        """segments: list of lists of parameters (requires_grad ones are kept)."""
        self.segments = [[p for p in seg if p.requires_grad] for seg in segments]
        params = [p for seg in self.segments for p in seg]
        if not params:
            raise ValueError('GradBucket: no trainable parameter')
        self.params = params
        self.flat = torch.zeros(sum(p.numel() for p in params), device=params[0].device, dtype=torch.float32)
        self.views, self.ranges = [], []
        o = 0
        for seg in self.segments:
            lo = o
            for p in seg:
                n = p.numel()
                self.views.append(self.flat[o:o + n].view_as(p))
                o += n
            self.ranges.append((lo, o))
        self.works = []
        self.attach()

    def attach(self):
        """(Re-)install the views as the parameters' `.grad` (optimizer.zero_grad(set_to_none=True) drops them)."""
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def zero(self):
        self.flat.zero_()
        self.attach()

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def reduce_async(self, k, group=None):
        """All-reduce (sum) segment k; returns immediately.  No-op without a process group / with one rank."""
        import torch.distributed as dist
        lo, hi = self.ranges[k]
        if hi == lo or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return 0
        self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
        return (hi - lo) * 4

    def reduce_all(self, group=None):
        return sum(self.reduce_async(k, group) for k in range(len(self.ranges)))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


_buckets = {}


def allreduce_gradients(params, group=None):
    """Sum the gradients over ranks IN PLACE through one persistent flat bucket (created on the first call for this
    parameter list; existing gradients are moved into it once).  Returns the number of bytes reduced.  No-op without
    an initialised process group."""
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    key = tuple(id(p) for p in params)
    old = [p.grad for p in params]
    b = _buckets.get(key)
    if b is None:
        _buckets.clear()                       # one parameter list at a time: drop the previous model's buffer
        b = _buckets[key] = GradBucket([params])
    b.attach()
    for v, g in zip(b.views, old):             # gradients autograd (or anyone) left outside the bucket move in once
        if g is None:
            v.zero_()
        elif g.data_ptr() != v.data_ptr():
            v.copy_(g)
    n = b.reduce_all(group)
    b.wait()
    return n


def _mm_tn(a2d, b2d):
    """a^T b (contraction over rows): tcgen05 bf16x3 when the tensor-core path is on, else the library GEMM."""
    if M.use_tensor_cores() and config.TRAIN_TC_GEMMS:
        return M.matmul_tn(a2d, b2d)
    return a2d.t() @ b2d


def _mm_nn(a2d, w):
    if M.use_tensor_cores() and config.TRAIN_TC_GEMMS:
        return M.matmul_nn(a2d, w)
    return a2d @ w


def _accum(p, g):
    g = g.reshape(p.shape).to(p.dtype)
    if p.grad is None:
        p.grad = g.clone()
    else:
        p.grad.add_(g)


class TrainStep(object):
    """loss_and_grads(mix_feas, spk_idx, target[, mix_mag]) -> loss tensors; parameter grads accumulated."""

    def __init__(self, mix_hidden_layer_3d, mix_speech_multiEmbedding, att_speech_layer, adjust_layer=None):
        if att_speech_layer.mode != 'dot':
            raise NotImplementedError("the training step implements the 'dot' attention")
        self.mix, self.emb, self.att = mix_hidden_layer_3d, mix_speech_multiEmbedding, att_speech_layer
        self.adj = adjust_layer if (adjust_layer is not None and config.is_SelfTune) else None
        self.complex_mask = bool(config.is_ComlexMask)
        self._bptt = {}           # (layer, B, T) -> static buffers + captured BPTT graph
        self._bucket = None       # GradBucket: created by step() on first use
        self._dz_planes = None    # bf16 hi/lo planes of d(loss)/d(pre-tanh embedding), reused across steps
        self._reduce = None       # (bucket, group) while a step() wants its segments reduced as they complete
        self.reduced_bytes = 0

    def segments(self):
        """Parameters in the order the backward pass completes their gradients: head (Linear, embedding table, ADDJUST),
        then the recurrent layers from the top one down."""
        rnn = self.mix.layer
        head = list(self.mix.Linear.parameters()) + list(self.emb.parameters())
        if self.adj is not None:
            head += list(self.adj.parameters())
        segs = [head]
        for l in range(rnn.num_layers - 1, -1, -1):
            segs.append([getattr(rnn, '%s_l%d%s' % (n, l, suf)) for suf in ('', '_reverse')
                         for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')])
        return segs

    def bucket(self):
        if self._bucket is None:
            self._bucket = GradBucket(self.segments())
        return self._bucket

    def _segment_done(self, k):
        if self._reduce is not None:
            b, group = self._reduce
            self.reduced_bytes += b.reduce_async(k, group)

    def parameters(self):
        mods = [self.mix, self.emb] + ([self.adj] if self.adj is not None else [])
        out = []
        for m in mods:
            out += list(m.parameters())
        return out

    # ------------------------------------------------------------------------------ forward
    def forward(self, mix_feas, spk_idx):
        """-> ctx dict with masks and everything backward needs."""
        lib = _lib.load()
        B, T, F = mix_feas.shape
        lin = self.mix.Linear
        E = lin.out_features // F
        saved = []
        bufs = None
        if config.TRAIN_CUDA_GRAPHS:       # static per-layer buffers so the BPTT chain can be replayed from a graph
            bufs = [self._bptt_state(l, B, T, mix_feas.device) for l in range(self.mix.layer.num_layers)]
        ex = {}
        hidden = M.rnn_forward(self.mix._packed, mix_feas, save=saved, buffers=bufs, extras=ex)   # K2 + K3, gates saved
        h_planes = ex.get('planes')            # the top layer's output as bf16 hi/lo planes [2, B*T, Kpy] (fused by K3)
        idx = self.emb.index_tensor(spk_idx)
        table = self.emb.layer.weight
        e = table.detach()[idx]                                                   # [B,S,EQ] gather (glue)
        S, EQ = e.shape[1], e.shape[2]
        hmean = None
        if self.adj is not None:
            hmean = hidden.mean(1)                                                # [B,2H]
            cat = torch.cat([hmean.unsqueeze(1).expand(B, S, hmean.shape[1]), e], 2)
            q = e + cat @ self.adj.layer.weight.detach().t()
        else:
            q = e
        q = q.contiguous()
        K = hidden.shape[2]
        h2d = hidden.view(B * T, K)
        if M.use_tensor_cores():
            if h_planes is not None and h_planes.shape[-1] > K:
                # a column of ones next to the K outputs: the weight-gradient GEMM over these planes then yields the bias
                # gradient as its last column (the forward product meets zero padding of the weight planes there)
                h_planes[0, :, K].fill_(1.0)
            emb = M.linear_tc(h_planes if h_planes is not None else M.split_bf16(h2d), M.weight_planes(lin.weight),
                              lin.bias.detach(), B * T, F * E, K, act='tanh')
        else:
            emb = M.linear_fwd(h2d, lin.weight.detach(), lin.bias.detach(), 'tanh')
        cplx = self.complex_mask
        mode = _lib.ATT_DOT_CRM if cplx else _lib.ATT_DOT
        masks = torch.empty((B, S, T, F, 2) if cplx else (B, S, T, F), device=hidden.device, dtype=torch.float32)
        rc = lib.dl4ss_attn_dot_fwd(_lib.ptr(emb), T * F * E, _lib.ptr(q), B, S, T * F, E, mode,
                                    float(config.cRM_k), float(config.cRM_C), _lib.ptr(masks), _lib.stream())
        _lib.check(rc, 'dl4ss_attn_dot_fwd')
        return {'B': B, 'T': T, 'F': F, 'E': E, 'S': S, 'EQ': EQ, 'saved': saved, 'hidden': hidden, 'idx': idx,
                'e': e, 'hmean': hmean, 'q': q, 'emb': emb, 'masks': masks, 'x0': mix_feas, 'h_planes': h_planes}

    # ------------------------------------------------------------------------------ loss + backward
    def loss_and_grads(self, mix_feas, spk_idx, target, mix_mag=None, global_batch=None, grad_scale=1.0, pit=False):
        """One forward + backward.  `global_batch`: utterances of the whole (all-rank) batch, so the
        MSELoss means -- and therefore the summed gradients -- are those of the global batch.
        Returns (loss, part0, part1) of THIS shard's contribution (sum over ranks = global loss)."""
        lib = _lib.load()
        with torch.no_grad():
            ctx = self.forward(mix_feas, spk_idx)
            B, T, F, E, S = ctx['B'], ctx['T'], ctx['F'], ctx['E'], ctx['S']
            Bg = B if global_batch is None else int(global_batch)
            cplx = self.complex_mask
            masks = ctx['masks']
            mix = mix_mag if cplx else mix_feas
            kind = _lib.MASK_COMPLEX if cplx else _lib.MASK_REAL
            if pit:     # permutation-invariant training: the targets are re-ordered per utterance to the best assignment
                from .pipeline import pit_mask_loss
                _, perms = pit_mask_loss(masks, mix, target, cplx)
                target = target[torch.arange(B, device=target.device)[:, None], perms].contiguous()
                ctx['perms'] = perms
            acc = torch.zeros(2, device=masks.device, dtype=torch.float64)
            rc = lib.dl4ss_mask_loss_fwd(_lib.ptr(masks), kind, _lib.ptr(mix, name='mix'), _lib.ptr(target, name='target'),
                                         B, S, T * F, _lib.ptr(acc, torch.float64), _lib.stream())
            _lib.check(rc, 'dl4ss_mask_loss_fwd')
            n0 = float(Bg * S * T * F)
            n1 = float(Bg * T * F)
            if cplx:
                l0, l1 = acc[0] / n0, acc[1] / n0
                loss = l0 + l1
                c0 = c1 = 2.0 * grad_scale / n0
            else:
                l0, l1 = acc[0] / n0, acc[1] / n1
                loss = l0 + 0.5 * l1
                c0, c1 = 2.0 * grad_scale / n0, grad_scale / n1
            dmask = torch.empty_like(masks)
            rc = lib.dl4ss_mask_loss_bwd(_lib.ptr(masks), kind, _lib.ptr(mix), _lib.ptr(target), B, S, T * F,
                                         c0, c1, _lib.ptr(dmask), _lib.stream())
            _lib.check(rc, 'dl4ss_mask_loss_bwd')
            self.backward(ctx, dmask)
        return loss, l0, l1

    def backward(self, ctx, dmask):
        lib = _lib.load()
        B, T, F, E, S, EQ = ctx['B'], ctx['T'], ctx['F'], ctx['E'], ctx['S'], ctx['EQ']
        cplx = self.complex_mask
        mode = _lib.ATT_DOT_CRM if cplx else _lib.ATT_DOT
        emb, q, masks, hidden = ctx['emb'], ctx['q'], ctx['masks'], ctx['hidden']
        dq = torch.empty_like(q)
        lin = self.mix.Linear
        h2d = hidden.view(B * T, -1)
        K = h2d.shape[1]
        h_planes = ctx.get('h_planes')
        if (M.use_tensor_cores() and config.TRAIN_TC_GEMMS and config.TRAIN_MN_GEMMS and h_planes is not None
                and h_planes.shape[-1] > K):
            # attention + tanh backward with dz emitted as bf16 planes: dW_lin (+ the bias gradient, through the ones column of
            # the h planes) from MN-major operands and dh from the same planes -- no fp32 dz, no split passes over it
            ldp = (F * E + 7) // 8 * 8
            dzp = self._dz_planes
            if dzp is None or dzp.shape != (2, B * T, ldp) or dzp.device != emb.device:
                dzp = torch.zeros(2, B * T, ldp, device=emb.device, dtype=torch.bfloat16)      # pad columns stay zero
                self._dz_planes = dzp
            rc = lib.dl4ss_attn_dot_bwd_planes(_lib.ptr(emb), _lib.ptr(q), _lib.ptr(masks), _lib.ptr(dmask), B, S, T, F, E, mode,
                                               float(config.cRM_k), float(config.cRM_C), _lib.ptr(dzp, torch.bfloat16), ldp,
                                               _lib.ptr(dq), _lib.stream())
            _lib.check(rc, 'dl4ss_attn_dot_bwd_planes')
            wext = M.linear_tc_tn(dzp, 0, F * E, 0, h_planes, 0, K + 1, 0, B, T, wa=F * E, wb=K + 1)   # [F*E, K+1]
            _accum(lin.weight, wext[:, :K])
            _accum(lin.bias, wext[:, K])
            dh = torch.empty(B * T, K, device=emb.device, dtype=torch.float32)
            rc = lib.dl4ss_linear_tc_lda_fwd(_lib.ptr(dzp, torch.bfloat16), ldp, _lib.ptr(M.weight_t_planes(lin.weight), torch.bfloat16),
                                             None, _lib.ptr(dh), K, B * T, K, F * E, _lib.stream())
            _lib.check(rc, 'dl4ss_linear_tc_lda_fwd')
            dh = dh.view(B, T, -1)
        else:
            # attention + tanh backward: emb is overwritten by dz
            rc = lib.dl4ss_attn_dot_bwd(_lib.ptr(emb), _lib.ptr(q), _lib.ptr(masks), _lib.ptr(dmask), B, S, T * F, E, mode,
                                        float(config.cRM_k), float(config.cRM_C), _lib.ptr(emb), _lib.ptr(dq), _lib.stream())
            _lib.check(rc, 'dl4ss_attn_dot_bwd')
            dz = emb.view(B * T, F * E)
            _accum(lin.weight, _mm_tn(dz, h2d))
            _accum(lin.bias, dz.sum(0))
            dh = _mm_nn(dz, lin.weight.detach()).view(B, T, -1)
        # speaker query backward (tiny: glue in torch)
        table = self.emb.layer.weight
        if self.adj is not None:
            W = self.adj.layer.weight.detach()                                   # [EQ, C+EQ]
            C = hidden.shape[2]
            cat = torch.cat([ctx['hmean'].unsqueeze(1).expand(B, S, C), ctx['e']], 2)
            _accum(self.adj.layer.weight, dq.reshape(B * S, EQ).t() @ cat.reshape(B * S, C + EQ))
            dcat = dq @ W                                                        # [B,S,C+EQ]
            de = dq + dcat[..., C:]
            dh = dh + (dcat[..., :C].sum(1) / float(T)).unsqueeze(1)
        else:
            de = dq
        gt = torch.zeros_like(table)
        gt.index_add_(0, ctx['idx'].reshape(-1), de.reshape(B * S, EQ))
        _accum(table, gt)
        self._segment_done(0)                  # head gradients complete: their all-reduce runs under the encoder backward
        # encoder backward, top layer first
        self.rnn_backward(ctx, dh.contiguous())

    def _bptt_state(self, l, B, T, dev):
        key = (l, B, T)
        st = self._bptt.get(key)
        if st is None:
            rnn = self.mix.layer
            G = 3 if isinstance(rnn, nn.GRU) else 4
            H = rnn.hidden_size
            f = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
            st = {'y': f(B, T, 2 * H), 'gates': f(B, T, 2, G * H), 'cells': f(B, T, 2, H), 'dy': f(B, T, 2 * H),
                  'dgx': f(B, T, 2, G * H), 'dgh': f(B, T, 2, G * H) if G == 3 else None, 'carry': f(2, B, H),
                  'dg_cur': f(2, B, G * H), 'dh_rec': f(2, B, H), 'whh': f(2, G * H, H), 'graph': None}
            self._bptt[key] = st
        return st

    def _bptt_chain(self, cell, st, B, T, H):
        """The T-step BPTT chain of one layer: per step one library GEMM (dh_rec = dgates x W_hh, both directions
        batched) and one gate kernel."""
        lib = _lib.load()
        for s in range(T):
            if s > 0:
                torch.bmm(st['dg_cur'], st['whh'], out=st['dh_rec'])
            rc = lib.dl4ss_rnn_bwd_step(cell, s, _lib.ptr(st['dy']), _lib.ptr(st['dh_rec']), _lib.ptr(st['gates']),
                                        _lib.ptr(st['cells']), _lib.ptr(st['y']), _lib.ptr(st['carry']),
                                        _lib.ptr(st['dgx']), _lib.ptr(st['dgh']), _lib.ptr(st['dg_cur']), B, T, H,
                                        _lib.stream())
            _lib.check(rc, 'dl4ss_rnn_bwd_step')

    def _bptt_persistent(self, cell, st, dy, whh, B, T, H, planes_only=False):
        """The T-step BPTT chain of one layer as one persistent kernel (csrc/rnn_bwd.cu).  planes_only (LSTM, tensor-core
        kernel): the gate gradients leave as bf16 planes only, no fp32 copy."""
        lib = _lib.load()
        need = lib.dl4ss_rnn_bwd_workspace_bytes(B, T, H, cell)
        ws = st.get('bwd_ws')
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, device=dy.device, dtype=torch.uint8)
            st['bwd_ws'] = ws
        if M.use_tensor_cores() and lib.dl4ss_rnn_bwd_tc_supported(H, cell) != 0:
            xb = lib.dl4ss_rnn_bwd_tc_xplanes_bytes(B, T, H, cell)
            xp = st.get('xplanes')
            if xp is None or xp.numel() != xb:
                xp = torch.zeros(xb, device=dy.device, dtype=torch.uint8)      # pad columns stay zero for good
                st['xplanes'] = xp
            rc = lib.dl4ss_rnn_layer_bwd_tc(cell, _lib.ptr(dy), _lib.ptr(whh), _lib.ptr(st['gates']),
                                            _lib.ptr(st['cells']), _lib.ptr(st['y']), _lib.ptr(None if planes_only else st['dgx']),
                                            _lib.ptr(st['dgh']), ctypes.c_void_p(xp.data_ptr()), B, T, H,
                                            ctypes.c_void_p(ws.data_ptr()), need, _lib.stream())
            _lib.check(rc, 'dl4ss_rnn_layer_bwd_tc')
            return
        rc = lib.dl4ss_rnn_layer_bwd(cell, _lib.ptr(dy), _lib.ptr(whh), _lib.ptr(st['gates']), _lib.ptr(st['cells']),
                                     _lib.ptr(st['y']), _lib.ptr(st['dgx']), _lib.ptr(st['dgh']), B, T, H,
                                     ctypes.c_void_p(ws.data_ptr()), need, _lib.stream())
        _lib.check(rc, 'dl4ss_rnn_layer_bwd')

    def rnn_backward(self, ctx, dy):
        rnn = self.mix.layer
        gru = isinstance(rnn, nn.GRU)
        cell = _lib.CELL_GRU if gru else _lib.CELL_LSTM
        G = 3 if gru else 4
        H = rnn.hidden_size
        B, T = ctx['B'], ctx['T']
        dev = dy.device
        layers = self.mix._packed.get()
        use_graph = bool(config.TRAIN_CUDA_GRAPHS)
        for l in range(rnn.num_layers - 1, -1, -1):
            sv, lw = ctx['saved'][l], layers[l]
            if use_graph:
                st = self._bptt_state(l, B, T, dev)          # y / gates / cells were written here by the forward
            else:
                st = {'y': sv['y'], 'gates': sv['gates'], 'cells': sv['cells'], 'dy': torch.empty_like(dy),
                      'dgx': torch.empty(B, T, 2, G * H, device=dev), 'dgh': torch.empty(B, T, 2, G * H, device=dev) if gru else None,
                      'carry': torch.empty(2, B, H, device=dev), 'dg_cur': torch.empty(2, B, G * H, device=dev),
                      'dh_rec': torch.empty(2, B, H, device=dev), 'whh': lw['whh'], 'graph': None}
            persistent = bool(config.TRAIN_PERSISTENT_BPTT) and _lib.load().dl4ss_rnn_bwd_supported(H, cell) != 0
            # LSTM on the MN-major path: every gradient of the layer (dW_ih, dW_hh, dX and -- through a column of ones next to the
            # layer input's planes -- the bias) is taken from the BPTT kernel's bf16 planes, so the fp32 copy is not written at all
            nin_l = sv['x'].shape[-1]
            planes_only = bool(
                persistent and not gru and M.use_tensor_cores() and config.TRAIN_TC_GEMMS and config.TRAIN_MN_GEMMS
                and getattr(config, 'TRAIN_PLANES_ONLY_BPTT', True)
                and _lib.load().dl4ss_rnn_bwd_tc_supported(H, cell) != 0 and (G * H) % 8 == 0
                and sv.get('x_planes') is not None and sv.get('y_planes') is not None and sv['x_planes'].shape[-1] > nin_l)
            if persistent:
                self._bptt_persistent(cell, st, dy, lw['whh'], B, T, H, planes_only)      # the whole chain in one launch
            elif use_graph and st['graph'] is not None:
                st['dy'].copy_(dy)
                st['whh'].copy_(lw['whh'])
                st['graph'].replay()
            else:
                st['dy'].copy_(dy)
                if use_graph:
                    st['whh'].copy_(lw['whh'])
                self._bptt_chain(cell, st, B, T, H)          # eager: also warms cuBLAS up before any capture
                if use_graph and T > 4:
                    # the chain is launch bound (2 launches per time step): capture it once, replay it from now on
                    cur = torch.cuda.current_stream()
                    side = torch.cuda.Stream()
                    side.wait_stream(cur)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.stream(side):
                        with torch.cuda.graph(graph, stream=side):
                            self._bptt_chain(cell, st, B, T, H)
                    cur.wait_stream(side)
                    st['graph'] = graph
            dgx, dgh = st['dgx'], st['dgh']
            x2d = sv['x'].reshape(B * T, -1)
            dgx2d = dgx.view(B * T, 2 * G * H)
            y = sv['y']
            db_x = None if planes_only else dgx2d.sum(0)
            if not gru and M.use_tensor_cores() and config.TRAIN_TC_GEMMS:
                # LSTM: the recurrent-side gate gradients ARE dgx, so one transposed split of dgx serves dW_ih and both
                # directions' dW_hh: the latter come out of ONE product against the time-shifted layer output
                # (columns [0,H): y[t-1] of the forward direction, [H,2H): y[t+1] of the reverse one; the rows without
                # a predecessor stay zero) as the two diagonal blocks.
                R = B * T
                xp = st.get('xplanes') if persistent else None
                GHg = (G * H + 7) // 8 * 8
                if (config.TRAIN_MN_GEMMS and xp is not None and sv.get('x_planes') is not None and sv.get('y_planes') is not None
                        and xp.numel() == 2 * R * 2 * GHg * 2):
                    # MN-major operands: the BPTT kernel's bf16 planes of the gate gradients, the forward pass's planes of the
                    # layer input and output are used as they lie (row-major): no transposing split, and the time shift of
                    # the recurrent products is a TMA coordinate offset (frames outside the utterance read zeros)
                    dg_pl = xp.view(torch.bfloat16).view(2, R, 2 * GHg)
                    x_pl, y_pl = sv['x_planes'], sv['y_planes']
                    nin = x2d.shape[1]
                    if planes_only:
                        # bias gradient = dgates^T 1: a column of ones behind the layer input's columns (hi plane; the padding of
                        # the lo plane is zero), one more output column of the same product
                        x_pl[0, :, nin] = 1.0
                        dW_ext = M.linear_tc_tn(dg_pl, 0, 2 * G * H, 0, x_pl, 0, nin + 1, 0, B, T, wa=2 * GHg, wb=nin + 1)
                        x_pl[0, :, nin] = 0.0
                        dW_ih, db_x = dW_ext[:, :nin], dW_ext[:, nin]
                    else:
                        dW_ih = M.linear_tc_tn(dg_pl, 0, 2 * G * H, 0, x_pl, 0, nin, 0, B, T, wa=2 * GHg, wb=nin) if GHg == G * H else None
                    for d, suf in enumerate(('', '_reverse')):
                        if dW_ih is None:
                            dWd = M.linear_tc_tn(dg_pl, d * GHg, G * H, 0, x_pl, 0, nin, 0, B, T, wa=2 * GHg, wb=nin)
                        else:
                            dWd = dW_ih[d * G * H:(d + 1) * G * H]
                        dWh = M.linear_tc_tn(dg_pl, d * GHg, G * H, 0, y_pl, d * H, H, -1 if d == 0 else 1, B, T,
                                             wa=2 * GHg, wb=2 * H)
                        sl = slice(d * G * H, (d + 1) * G * H)
                        _accum(getattr(rnn, 'weight_ih_l%d%s' % (l, suf)), dWd)
                        _accum(getattr(rnn, 'bias_ih_l%d%s' % (l, suf)), db_x[sl])
                        _accum(getattr(rnn, 'weight_hh_l%d%s' % (l, suf)), dWh)
                        _accum(getattr(rnn, 'bias_hh_l%d%s' % (l, suf)), db_x[sl])
                else:
                    dg_t_planes = M.split_bf16_t(dgx2d)
                    dW_ih = M.linear_tc(dg_t_planes, M.split_bf16_t(x2d), None, 2 * G * H, x2d.shape[1], R, split_k=True)
                    ysh = st.get('ysh')
                    if ysh is None or ysh.shape != y.shape:
                        ysh = torch.zeros_like(y)
                        st['ysh'] = ysh
                    ysh[:, 1:, :H].copy_(y[:, :-1, :H])
                    ysh[:, :-1, H:].copy_(y[:, 1:, H:])
                    P = M.linear_tc(dg_t_planes, M.split_bf16_t(ysh.view(R, 2 * H)), None, 2 * G * H, 2 * H, R, split_k=True)
                    for d, suf in enumerate(('', '_reverse')):
                        sl = slice(d * G * H, (d + 1) * G * H)
                        _accum(getattr(rnn, 'weight_ih_l%d%s' % (l, suf)), dW_ih[sl])
                        _accum(getattr(rnn, 'bias_ih_l%d%s' % (l, suf)), db_x[sl])
                        _accum(getattr(rnn, 'weight_hh_l%d%s' % (l, suf)), P[sl, d * H:(d + 1) * H])
                        _accum(getattr(rnn, 'bias_hh_l%d%s' % (l, suf)), db_x[sl])
            else:
                dW_ih = _mm_tn(dgx2d, x2d)                                       # [2*G*H, in]
                dgr = dgh if gru else dgx                                        # recurrent-side gate grads
                for d, suf in enumerate(('', '_reverse')):
                    sl = slice(d * G * H, (d + 1) * G * H)
                    _accum(getattr(rnn, 'weight_ih_l%d%s' % (l, suf)), dW_ih[sl])
                    _accum(getattr(rnn, 'bias_ih_l%d%s' % (l, suf)), db_x[sl])
                    if d == 0:       # h_{t-1} of the forward direction is y[:, t-1, :H]
                        dg_t, h_prev = dgr[:, 1:, 0, :], y[:, :-1, :H]
                    else:            # the reverse direction came from t+1
                        dg_t, h_prev = dgr[:, :-1, 1, :], y[:, 1:, H:]
                    _accum(getattr(rnn, 'weight_hh_l%d%s' % (l, suf)),
                           _mm_tn(dg_t.reshape(-1, G * H), h_prev.reshape(-1, H)))
                    _accum(getattr(rnn, 'bias_hh_l%d%s' % (l, suf)), dgr[:, :, d, :].sum((0, 1)))
            self._segment_done(rnn.num_layers - l)      # this layer's gradients: reduced under the layers below
            if l > 0:
                xp = st.get('xplanes') if (persistent and not gru) else None
                GHg = (G * H + 7) // 8 * 8
                if (M.use_tensor_cores() and config.TRAIN_TC_GEMMS and config.TRAIN_MN_GEMMS and xp is not None and GHg == G * H
                        and xp.numel() == 2 * B * T * 2 * GHg * 2):
                    # dX = dgates W_ih from the BPTT kernel's planes as they lie (row pitch 2*G*H): no split of dgx
                    nin = lw['wih'].shape[1]
                    out = torch.empty(B * T, nin, device=dev, dtype=torch.float32)
                    rc = _lib.load().dl4ss_linear_tc_lda_fwd(ctypes.c_void_p(xp.data_ptr()), 2 * GHg,
                                                             _lib.ptr(M.weight_t_planes(lw['wih']), torch.bfloat16), None,
                                                             _lib.ptr(out), nin, B * T, nin, 2 * G * H, _lib.stream())
                    _lib.check(rc, 'dl4ss_linear_tc_lda_fwd')
                    dy = out.view(B, T, -1)
                else:
                    dy = _mm_nn(dgx2d, lw['wih']).view(B, T, -1).contiguous()

    # ------------------------------------------------------------------------------ full step
    def step(self, optimizer, mix_feas, spk_idx, target, mix_mag=None, global_batch=None, group=None, overlap=True,
             reduce=True):
        """zero the gradient bucket -> loss_and_grads -> all-reduce (if a process group is up; per segment, overlapped
        with the rest of the backward pass unless overlap=False; reduce=False skips the collective: timing only) ->
        optimizer.step.  `global_batch` defaults to the
        sum of the shard sizes over the ranks."""
        if global_batch is None:
            global_batch = global_batch_size(mix_feas.shape[0], mix_feas.device, group)
        b = self.bucket()
        b.zero()
        self.reduced_bytes = 0
        self._reduce = (b, group) if (overlap and reduce) else None
        try:
            out = self.loss_and_grads(mix_feas, spk_idx, target, mix_mag, global_batch)
        finally:
            self._reduce = None
        if reduce and not overlap:
            self.reduced_bytes = b.reduce_all(group)
        b.wait()
        optimizer.step()
        return out

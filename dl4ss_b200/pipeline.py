"""Waveform -> separated waveforms, the whole hot path on the GPU.

`Separator` strings the kernels in the order the reference's `eval_bss` does
(TDAA_beta/main_run_sstune_EvalVer.py:407-507, cRM: TDAA_beta/main_run_sstune_cRM_EvalVer.py:498-570):
    prepare_data (CPU librosa STFT)      -> K1  dl4ss_stft_feat          (one launch per batch)
    MIX_SPEECH (cuDNN RNN + Linear+tanh) -> per layer: input projection GEMM + K3 persistent RNN
    SPEECH_EMBEDDING + ADDJUST           -> dl4ss_speaker_query_fwd
    expand().contiguous() + ATTENTION    -> K4  dl4ss_emb_attn_mask_fwd   (no [B,T,F,E] tensor)
    mask*mix, bss_eval (CPU librosa istft, wav files) -> K6 dl4ss_mask_istft
It takes the drop-in modules (dl4ss_b200.modules), so weights/checkpoints are the reference's.
"""
import torch

from . import _lib
from . import config
from . import features
from . import modules as M


class Separator(object):
    def __init__(self, mix_hidden_layer_3d, mix_speech_multiEmbedding, att_speech_layer, adjust_layer=None,
                 n_fft=None, hop=None):
        if att_speech_layer.mode != 'dot':
            raise NotImplementedError("Separator fuses the 'dot' attention; run 'align' through the modules")
        self.mix = mix_hidden_layer_3d
        self.emb = mix_speech_multiEmbedding
        self.att = att_speech_layer
        self.adj = adjust_layer if (adjust_layer is not None and config.is_SelfTune) else None
        self.n_fft = config.FRAME_LENGTH if n_fft is None else n_fft
        self.hop = config.FRAME_SHIFT if hop is None else hop
        self.complex_mask = bool(config.is_ComlexMask)
        self.log_spectral = bool(config.IS_LOG_SPECTRAL)

    # -- stages ---------------------------------------------------------------------------
    def features(self, mix_wav):
        return features.prepare_batch(mix_wav, self.n_fft, self.hop, self.log_spectral)

    def queries(self, hidden, spk_idx, hmean=None):
        """Speaker queries [B,S,E|2E]; `hmean` [B,2H] (the encoder kernel's fused T-mean) replaces the pass over
        `hidden` that ADDJUST's mean needs."""
        idx = self.emb.index_tensor(spk_idx)
        wadj = self.adj.layer.weight.detach() if self.adj is not None else None
        h = None
        if wadj is not None:
            h = hmean.unsqueeze(1) if hmean is not None else hidden
        q, err = M._speaker_query(h, self.emb.layer.weight.detach(), idx, wadj, 1)
        return q, err

    def masks(self, mix_feas, spk_idx, check_index=True):
        """mix_feas [B,T,F], spk_idx int [B,S] -> masks [B,S,T,F] (cRM: decompressed [B,S,T,F,2])."""
        extras = {}
        lin = self.mix.Linear
        S = spk_idx.shape[1] if hasattr(spk_idx, 'shape') else len(spk_idx[0])
        # the fused Linear/tanh/attention kernel (E = 50, S <= 4) reads the encoder output as bf16 planes and ADDJUST its fused
        # T-mean: the fp32 copy of the hidden states is then never read and the recurrent kernel does not write it
        fused_head = M.use_tensor_cores() and lin.out_features // self.mix.input_fre == 50 and S <= 4
        hidden = self.mix.encode(mix_feas, extras, planes_only=fused_head)
        return self.masks_from_hidden(hidden, extras, spk_idx, check_index)

    def masks_from_hidden(self, hidden, extras, spk_idx, check_index=True):
        """The part of `masks` after the encoder: speaker queries + the fused Linear/tanh/attention kernel."""
        q, err = self.queries(hidden, spk_idx, extras.get('hmean'))
        lin = self.mix.Linear
        F = self.mix.input_fre
        E = lin.out_features // F
        out = M.emb_attn_mask(hidden, lin.weight, lin.bias, q, F, E,
                              complex_mask=self.complex_mask, decompress=True, h_planes=extras.get('planes'))
        self.last_err = err          # device flag of the embedding gather (GraphedSeparator.check_index reads it)
        if check_index and int(err.item()):
            raise IndexError('index out of range in self')
        return out

    def masks_multihot(self, mix_feas, top_k_mask_mixspeech):
        """The older Torch_multi form (Torch_multi/main_run_multi_selfSS.py:476-493): one mask per candidate speaker,
        `multi_mask * top_k_mask` -- channels the 0/1 selection `top_k_mask_mixspeech` [B,num_labels] switches off are
        exactly zero, so only the active channels are evaluated (by the same fused kernel, ADDJUST not applied: that
        script has none) and scattered into the [B,num_labels,T,F] result."""
        sel = top_k_mask_mixspeech.to(device=mix_feas.device, dtype=torch.float32)
        B, N = sel.shape
        T, F = mix_feas.shape[1:]
        count = sel.sum(1).to(torch.int64)
        S = max(int(count.max().item()), 1)
        order = torch.argsort(sel, dim=1, descending=True, stable=True)[:, :S]          # active channels first, by index
        valid = torch.arange(S, device=sel.device)[None, :] < count[:, None]
        idx = torch.where(valid, order, torch.zeros_like(order)).contiguous()
        adj, self.adj = self.adj, None
        try:
            extras = {}
            hidden = self.mix.encode(mix_feas, extras)                              # once; the fused kernel takes S <= 4 queries
            m = torch.cat([self.masks_from_hidden(hidden, extras, idx[:, k:k + 4].contiguous()) for k in range(0, S, 4)], 1)
        finally:
            self.adj = adj
        m = m * valid.to(m.dtype).view(B, S, *([1] * (m.dim() - 2)))
        out = torch.zeros((B, N) + tuple(m.shape[2:]), device=m.device, dtype=m.dtype)
        out.scatter_add_(1, idx.view(B, S, *([1] * (m.dim() - 2))).expand_as(m), m)
        return out

    def separate(self, mix_wav, spk_idx, return_all=False, check_index=True):
        """mix_wav [B,L] CUDA f32/f64, spk_idx [B,S] -> separated wav [B,S,hop*(T-1)] float32."""
        with torch.no_grad():
            batch = self.features(mix_wav)
            masks = self.masks(batch['mix_feas'], spk_idx, check_index)
            wav = features.mask_istft(masks, batch['mix_mag'], self.hop, 'hann', self.n_fft)
        if return_all:
            batch['masks'] = masks
            batch['wav'] = wav
            return batch
        return wav

    __call__ = separate

    def recursive_extract(self, mix_feas, classifier, steps=3, alpha=-0.5):
        """Recursive extract-and-subtract inference (TDAA_beta/main_run_sstune_RecuVer.py:32-79,480-494; the
        reference runs it at batch size 1 through numpy): per step `classifier` (MIX_SPEECH_classifier) +
        `top_k_mask(., alpha, 1)` name the most probable remaining speaker of every utterance, the mask path
        predicts that speaker's spectrogram from the current features, and the prediction is subtracted from the
        features.  Everything stays on the device, the whole batch advances together.
        mix_feas [B,T,F] -> (predict_multi_map [B,steps,T,F], speakers int64 [B,steps])."""
        if self.complex_mask:
            raise NotImplementedError('the recursive variant exists for real masks only (RecuVer.py)')
        with torch.no_grad():
            now = mix_feas.contiguous().clone()
            B, T, F = now.shape
            preds = torch.empty(B, steps, T, F, device=now.device, dtype=torch.float32)
            spk = torch.empty(B, steps, device=now.device, dtype=torch.int64)
            for k in range(steps):
                prob = classifier(now)
                sel = M.top_k_mask(prob, alpha, 1)                 # exactly one 1 per row when alpha < 0
                idx = sel.argmax(1, keepdim=True)
                masks = self.masks(now, idx, check_index=False)    # [B,1,T,F]
                torch.mul(masks[:, 0], now, out=preds[:, k])
                spk[:, k] = idx[:, 0]
                now = now - preds[:, k]
        return preds, spk


_ENV_TWO_CTA = int(__import__('os').environ.get('DL4SS_GEMM_2CTA', '1'))          # the library's own defaults (A/B switches)
_ENV_PAIRS = int(__import__('os').environ.get('DL4SS_RNN_CLUSTER_PAIRS', '0'))


class sm_sharing(object):
    """Launch geometry for batches that run side by side on different streams: the recurrent kernel packs `tiles`
    utterance tiles into every CTA (3 -> 60 CTAs for 256 utterances instead of 80) and a projection launch takes at
    most `gemm_ctas` SMs, so that the recurrent launch of one batch and a projection / recurrent launch of the other
    are co-resident.  The settings are process-wide in the library and read at launch time (a CUDA graph keeps the
    geometry it was captured with); (0, 0) = a lone batch owns the GPU."""
    PIPELINED = tuple(int(v) for v in __import__('os').environ.get('DL4SS_SM_SHARING', '3,88').split(','))   # (tiles per recurrent CTA, projection CTA cap)

    @staticmethod
    def for_batch(B, inflight):
        """Geometry for `inflight` batches of B utterances: the shared form while two recurrent launches fit side by side
        (B <= 288: 3 tile groups x 2 directions x 10 slices = 60 CTAs each), otherwise every launch takes what it needs."""
        return sm_sharing.PIPELINED if (inflight > 1 and B <= 288) else None

    def __init__(self, tiles, gemm_ctas):
        self.want = (int(tiles), int(gemm_ctas))

    def __enter__(self):
        lib = _lib.load()
        shared = self.want != (0, 0)
        lib.dl4ss_rnn_tc_set_tiles_per_cta(self.want[0])
        lib.dl4ss_gemm_tc_set_max_ctas(self.want[1])
        # shared geometry: the recurrent launch as 2-CTA clusters (it then fills whole TPCs) and the plain projections on
        # 2-CTA tiles under the cap as well (they need both SMs of a TPC): 6.70 -> 6.60 ms per step
        if shared:
            lib.dl4ss_rnn_tc_set_cluster_pairs(1)
            lib.dl4ss_gemm_tc_set_two_cta(2 if _ENV_TWO_CTA else 0)
        return self

    def __exit__(self, *exc):
        lib = _lib.load()
        lib.dl4ss_rnn_tc_set_tiles_per_cta(0)
        lib.dl4ss_gemm_tc_set_max_ctas(0)
        lib.dl4ss_rnn_tc_set_cluster_pairs(_ENV_PAIRS)
        lib.dl4ss_gemm_tc_set_two_cta(_ENV_TWO_CTA)
        return False


class GraphedSeparator(object):
    """The whole waveform -> separated-waveforms step of a `Separator` for one fixed (B, L, S), captured ONCE in a
    CUDA graph: the ~130 kernel launches of a step become a single cudaGraphLaunch, so the GPU never waits for
    the host (python + ctypes per launch) and host scheduling jitter cannot stretch a step.

        gs = GraphedSeparator(sep, B, L, S)
        out = gs(wav, idx)          # copies into the static inputs, replays, returns the static output `gs.out`
    `gs.wav` / `gs.idx` / `gs.out` are static device tensors: fill the inputs in place and call `gs.replay()` to
    skip the device-to-device copies; consume `gs.out` before the next replay (and re-read `gs.out` after it: a
    re-capture replaces the tensor).  Speaker indices are not range checked inside the graph; `gs.check_index()`
    reads the gather kernel's error flag of the last replay.  The weights may change between replays (optimizer
    step, load_state_dict): the graph is captured again when they do."""

    def __init__(self, separator, B, L, S, wav_dtype=torch.float32, device=None, share=None):
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.sep = separator
        self.dev = dev
        self.share = share          # (tiles per recurrent CTA, projection CTA cap): launch geometry for pipelined batches
        self.wav = torch.zeros(B, L, device=dev, dtype=wav_dtype)
        self.idx = torch.zeros(B, S, device=dev, dtype=torch.int64)
        self.captures = 0
        self._capture()

    def _param_key(self):
        """(storage address, in-place version) of every parameter the captured kernels read.  The graph bakes in the
        device pointers of operands DERIVED from the weights (bf16 planes, packed W_hh, concatenated W_ih): after an
        optimizer step or load_state_dict those are stale, so `replay` re-captures when this key changes."""
        sep = self.sep
        mods = [sep.mix, sep.emb] + ([sep.adj] if sep.adj is not None else [])
        return tuple((p.data_ptr(), p._version) for m in mods for p in m.parameters())

    def _capture(self):
        separator, dev = self.sep, self.dev
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # first-use work (weight planes, twiddles, workspaces) stays outside the graph
            separator.separate(self.wav, self.idx, check_index=False)
            separator.separate(self.wav, self.idx, check_index=False)
        side.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with sm_sharing(*(self.share or (0, 0))):      # grid sizes are baked into the captured kernel nodes
            with torch.cuda.graph(self.graph, stream=side):
                self.out = separator.separate(self.wav, self.idx, check_index=False)
                self.err = separator.last_err
        cur.wait_stream(side)
        self.key = self._param_key()
        self.captures += 1

    def replay(self):
        if self._param_key() != self.key:      # weights changed since the capture (training between evaluations)
            torch.cuda.current_stream(self.dev).synchronize()
            self._capture()
        self.graph.replay()
        return self.out

    def __call__(self, mix_wav, spk_idx):
        self.wav.copy_(mix_wav, non_blocking=True)
        self.idx.copy_(spk_idx, non_blocking=True)
        return self.replay()

    def check_index(self):
        if int(self.err.item()):
            raise IndexError('index out of range in self')


class PipelinedSeparator(object):
    """`depth` batches of one fixed (B, L, S) in flight on `depth` streams, one `GraphedSeparator` each, captured with
    the `sm_sharing.PIPELINED` geometry.  The recurrent layers are latency chains that leave most of every SM idle, so a
    second batch in flight fills the machine: 256 utterances take 60 CTAs, the recurrent launches of two batches run
    side by side and one batch's projections overlap the other's chain (B200, 256 x 5 s: 9.7 -> 7.5 ms per batch).

        pipe = PipelinedSeparator(sep, B, L, S)
        k = pipe.submit(wav, idx)       # device tensors; returns the slot; the caller's stream order is respected
        out = pipe.result(k)            # static output of slot k, valid on the current stream until slot k is resubmitted
    """

    def __init__(self, separator, B, L, S, depth=2, device=None, wav_dtype=torch.float32):
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.dev, self.depth = dev, depth
        share = sm_sharing.for_batch(B, depth)
        self.gs = [GraphedSeparator(separator, B, L, S, wav_dtype, dev, share=share) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(dev) for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_read = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0

    def submit(self, mix_wav, spk_idx):
        k = self.count % self.depth
        self.count += 1
        cur = torch.cuda.current_stream(self.dev)
        st = self.streams[k]
        st.wait_stream(cur)                       # inputs produced on the caller's stream
        st.wait_event(self.ev_read[k])            # the previous result of this slot has been consumed (see `result`)
        with torch.cuda.stream(st):
            self.gs[k](mix_wav, spk_idx)
            self.ev_done[k].record(st)
        return k

    def result(self, k):
        """Static output tensor of slot k; the current stream waits for the slot's kernels.  The NEXT submit to this slot
        waits for everything the current stream has enqueued up to the following `release(k)` (or `result` of it)."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ev_done[k])
        return self.gs[k].out

    def release(self, k):
        """The current stream is done reading slot k's output: the slot may be overwritten by its next submit."""
        self.ev_read[k].record(torch.cuda.current_stream(self.dev))

    def drain(self):
        cur = torch.cuda.current_stream(self.dev)
        for e in self.ev_done:
            cur.wait_event(e)


class HostPipeline(object):
    """Waveforms in pinned HOST memory -> separated waveforms in pinned HOST memory, with the H2D copy of
    batch i+1 and the D2H copy of batch i-1 overlapped with the kernels of batch i (three streams, `depth`
    device slots).  This is the call a user of the reference's eval loop makes per batch: the reference
    moves features to the GPU and predictions back once per batch as well
    (TDAA_beta/main_run_sstune_EvalVer.py:420 `.cuda()`, :60-61 `.data.cpu().numpy()`), serially.

        pipe = HostPipeline(sep, B, L, S)
        for h_wav, h_idx, h_out in batches: pipe.submit(h_wav, h_idx, h_out)
        pipe.drain()            # all h_out buffers are complete after this
    Each in-flight step needs its own h_out buffer (rotate >= depth of them).  With `graphs=True` (default) every
    device slot owns a `GraphedSeparator`: a step is two async copies and one graph launch."""

    def __init__(self, separator, B, L, S, depth=2, device=None, graphs=True, concurrent=True):
        self.sep = separator
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.depth = depth
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # with graphs every slot computes on its own stream with the shared-SM launch geometry: two batches in flight
        share = sm_sharing.for_batch(B, depth) if (graphs and concurrent) else None
        # two compute streams whatever the depth: a third recurrent launch would not fit next to two (3 x 60 CTAs) and its
        # half-placed groups would spin on SMs the projections need; further slots only decouple the copies
        self.s_cmp = [torch.cuda.Stream(dev) for _ in range(2)] if (graphs and concurrent and depth > 1) else None
        self.gs = [GraphedSeparator(separator, B, L, S, device=dev, share=share) for _ in range(depth)] if graphs else None
        if graphs:
            self.d_wav = [g.wav for g in self.gs]
            self.d_idx = [g.idx for g in self.gs]
        else:
            self.d_wav = [torch.empty(B, L, device=dev, dtype=torch.float32) for _ in range(depth)]
            self.d_idx = [torch.empty(B, S, device=dev, dtype=torch.int64) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_copied = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = torch.cuda.Event()
        self.count = 0

    def submit(self, h_wav, h_idx, h_out):
        k = self.count % self.depth
        self.count += 1
        cur = self.s_cmp[(self.count - 1) % 2] if self.s_cmp is not None else torch.cuda.current_stream()
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_done[k])          # slot k's previous kernels have consumed d_wav[k]
            self.d_wav[k].copy_(h_wav, non_blocking=True)
            self.d_idx[k].copy_(h_idx, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        cur.wait_event(self.ev_in[k])
        with torch.cuda.stream(cur):
            if self.gs is not None:
                cur.wait_event(self.ev_copied[k])          # slot k's static output has left for the host
                out = self.gs[k].replay()
            else:
                out = self.sep.separate(self.d_wav[k], self.d_idx[k], check_index=False)
            self.ev_done[k].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_done[k])
            h_out.copy_(out, non_blocking=True)
            if self.gs is None:
                out.record_stream(self.s_out)
            self.ev_copied[k].record(self.s_out)
            self.ev_out.record(self.s_out)
        return h_out

    def drain(self):
        """Make the current stream wait for every submitted D2H copy (then synchronize it to read h_out)."""
        torch.cuda.current_stream().wait_event(self.ev_out)


def mask_loss(masks, mix, target, complex_mask=None):
    """The reference's training/eval objective from masks (K5).

    real : MSE(mask*mix_feas, y) + 0.5*MSE(sum_s mask, 1)  (TDAA_beta/main_run_sstune_EvalVer.py:487-497)
           mix = mix_feas [B,T,F], target [B,S,T,F]
    cRM  : MSE(Re) + MSE(Im)                              (...cRM_EvalVer.py:545-568)
           mix = mix_mag [B,T,F,2], target [B,S,T,F,2]
    -> (loss, part0, part1) python floats... as 0-d CUDA float64 tensors (no host sync here)."""
    lib = _lib.load()
    cplx = (masks.dim() == 5) if complex_mask is None else complex_mask
    B, S, T, F = masks.shape[:4]
    acc = torch.zeros(2, device=masks.device, dtype=torch.float64)
    rc = lib.dl4ss_mask_loss_fwd(_lib.ptr(masks, name='masks'), _lib.MASK_COMPLEX if cplx else _lib.MASK_REAL,
                                 _lib.ptr(mix, name='mix'), _lib.ptr(target, name='target'), B, S, T * F,
                                 _lib.ptr(acc, torch.float64), _lib.stream())
    _lib.check(rc, 'dl4ss_mask_loss_fwd')
    if cplx:
        n = float(B * S * T * F)
        l0, l1 = acc[0] / n, acc[1] / n
        return l0 + l1, l0, l1
    l0 = acc[0] / float(B * S * T * F)
    l1 = acc[1] / float(B * T * F)
    return l0 + 0.5 * l1, l0, l1


def pit_mask_loss(masks, mix, target, complex_mask=None):
    """Permutation-invariant MSE of mask x mixture against the targets (the north-star's PIT form of K5; the reference
    pairs sources by sorted speaker index, TDAA_beta/main_run_sstune_EvalVer.py:632-639).

    One kernel pass builds the S x S squared-error sums of every utterance (dl4ss_mask_pair_loss_fwd); the S! permutations
    are searched on the device.  Returns (loss, perms): loss = mean over utterances of the best assignment's MSE (cRM:
    MSE(Re) + MSE(Im)) as a 0-d CUDA float64 tensor, perms int64 [B,S] with perms[b,s] = the target matched to prediction s
    (oracle: oracle/modules_ref.py pit_mse_ref)."""
    import itertools
    lib = _lib.load()
    cplx = (masks.dim() == 5) if complex_mask is None else complex_mask
    B, S, T, F = masks.shape[:4]
    pair = torch.empty(B, S, S, device=masks.device, dtype=torch.float64)
    rc = lib.dl4ss_mask_pair_loss_fwd(_lib.ptr(masks, name='masks'), _lib.MASK_COMPLEX if cplx else _lib.MASK_REAL,
                                      _lib.ptr(mix, name='mix'), _lib.ptr(target, name='target'), B, S, T * F,
                                      _lib.ptr(pair, torch.float64), _lib.stream())
    _lib.check(rc, 'dl4ss_mask_pair_loss_fwd')
    perms = torch.tensor(list(itertools.permutations(range(S))), device=masks.device)          # [P,S]
    cost = pair[:, torch.arange(S, device=masks.device), perms].sum(-1) / float(S * T * F)          # [B,P]
    best = cost.argmin(1)
    return cost.gather(1, best[:, None]).mean(), perms[best]

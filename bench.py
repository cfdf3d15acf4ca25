#!/usr/bin/env python
"""bench.py -- separated audio-seconds per second of the DL4SS hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A step = one pass of the hot path (waveform -> STFT features -> BLSTM 4x300 -> speaker attention
masks -> mask x mixture -> iSTFT) over one batch of synthetic WSJ0-2mix-shaped mixtures
(BASELINE.json configs[1]: TDAA_beta, 2 speakers, 5 s @ 8 kHz, batch 256 per GPU, fp32).
`value` times the steps with the inputs resident in HBM; `e2e` times the same call from pinned host
buffers (H2D of the waveforms, D2H of the separated waveforms inside the timed region).
N > 1: one process per GPU (torchrun), utterances sharded by batch, no data-path collective
(inference), weak scaling.  `--impl reference` times the CPU oracle (a Python-3 restatement of the
reference, which is Python 2 + librosa and cannot run here) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

SR = 8000
WORKLOAD = {'name': 'TDAA_beta 2-speaker attention-mask inference (BASELINE configs[1])',
            'L': 40000, 'S': 2, 'hop': 128, 'n_fft': 256, 'cell': 'lstm', 'layers': 4, 'H': 300, 'E': 50,
            'num_spk': 101, 'self_tune': True, 'complex_mask': False}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d.get('hbm_gbs', 6650.0), 'bf16_tflops': d.get('bf16_tflops', 1590.0),
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', 1400.0), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the newest committed
    `ncu --set full` summary under profiles/ (None when there is none)."""
    import glob
    import re
    best = None
    def tag(path):          # full_<what>_r<round><letter>.md: newest round / letter last
        m = re.search(r'_r(\d+)([a-z]*)\.md$', path)
        return (int(m.group(1)), m.group(2), path) if m else (0, '', path)
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'full_*.md')), key=tag):
        txt = open(path).read()
        for sec in txt.split('### ')[1:]:
            if kernel_substr not in sec.split('\n', 1)[0]:
                continue
            rd = re.search(r'dram__bytes_read\.sum \| ([0-9.]+) \| (\w+)', sec)
            wr = re.search(r'dram__bytes_write\.sum \| ([0-9.]+) \| (\w+)', sec)
            if rd and wr:
                mul = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
                best = {'bytes_per_launch': float(rd.group(1)) * mul[rd.group(2)] + float(wr.group(1)) * mul[wr.group(2)],
                        'source': os.path.relpath(path, ROOT)}
                break
    return best


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Launch nvidia-smi in loop mode.  Called BEFORE the warm-up steps: its start-up (NVML init) can stall the
        driver for milliseconds, which must not land inside the timed region."""
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout:
            time.sleep(0.01)

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        if self.t1 is None:
            self.t1 = time.time()
        time.sleep(0.08)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = self.rows
        if self.t0 is not None:      # samples taken while the timed region ran (the one just after it if it was short)
            inside = [r for r in rows if self.t0 <= r[0] <= self.t1 + 0.06]
            rows = inside or rows[-1:]
        for _, r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_wave_batches(B, L, n_batches, device, seed):
    """Synthetic mixtures shaped like the reference's: S sources, each -mean, /max|.|, +-2.5 dB gain,
    summed (TDAA_beta/predata_fromList.py:146-177).  Built on the device (timing is data independent);
    the parity tests use the speech-like generator of oracle/synth.py."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(n_batches):
        mix = torch.zeros(B, L, device=device)
        for s in range(WORKLOAD['S']):
            x = torch.randn(B, L, device=device, generator=g)
            x = x - x.mean(1, keepdim=True)
            x = x / x.abs().amax(1, keepdim=True)
            gain = 10.0 ** ((torch.rand(B, 1, device=device, generator=g) * 5.0 - 2.5) / 20.0)
            mix += gain * x
        out.append(mix.contiguous())
    return out


def build_model(device):
    import dl4ss_b200 as d
    W = WORKLOAD
    torch.manual_seed(1)
    d.config.HIDDEN_UNITS, d.config.EMBEDDING_SIZE = W['H'], W['E']
    d.config.is_ComlexMask, d.config.is_SelfTune = W['complex_mask'], W['self_tune']
    d.config.FRAME_SHIFT = W['hop']
    T = 1 + W['L'] // W['hop']
    F = W['n_fft'] // 2 + 1
    mix = d.MIX_SPEECH(F, T, cell=W['cell'], num_layers=W['layers']).to(device)
    emb = d.SPEECH_EMBEDDING(W['num_spk'], W['E'], 2).to(device)
    att = d.ATTENTION(W['E'], 'dot').to(device)
    adj = d.ADDJUST(2 * W['H'], W['E']).to(device)
    return d.Separator(mix, emb, att, adj, W['n_fft'], W['hop'])


def kernel_stage_times(sep, wavs, idx, inner=8):
    """STFT and mask+iSTFT kernel times without host launch overhead: `inner` launches over rotating input
    batches (4 x 41 MB of waveforms / 4 x 82 MB of spectra: working set > 126 MB L2) captured in a CUDA graph
    and replayed; CUDA events on the replaying stream.  These are the HBM-bound stages whose roofline the
    north star asks for; their kernels run ~50-100 us, the same order as a Python-side launch."""
    from dl4ss_b200 import features
    W = WORKLOAD
    ev = lambda: torch.cuda.Event(enable_timing=True)
    outs = {}
    with torch.no_grad():
        batches = [sep.features(w) for w in wavs]
        B, T, F = batches[0]['mix_feas'].shape
        masks = torch.rand(B, W['S'], T, F, device=wavs[0].device)
        o_feat = [torch.empty_like(batches[0]['mix_feas']) for _ in range(2)]
        o_cplx = [torch.empty_like(batches[0]['mix_mag']) for _ in range(2)]
        o_wav = [torch.empty(B, W['S'], W['hop'] * (T - 1), device=wavs[0].device) for _ in range(2)]
        fns = {'stft': lambda i: features.stft_features(wavs[i % len(wavs)], W['n_fft'], W['hop'], 'hann', 'abs',
                                                        out_feat=o_feat[i % 2], out_cplx=o_cplx[i % 2]),
               'mask_istft': lambda i: features.mask_istft(masks, batches[i % len(batches)]['mix_mag'], W['hop'],
                                                           out=o_wav[i % 2])}
        for name, fn in fns.items():
            for i in range(2):
                fn(i)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    for i in range(inner):
                        fn(i)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = ev(), ev(); e0.record()
                graph.replay()
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / inner)
            outs[name] = best
    return outs


def stage_times(sep, wav, idx, reps=3):
    """Per-stage device time (CUDA events on the launching stream) of one step, for the roofline
    lines; run outside the timed region."""
    import dl4ss_b200 as d
    from dl4ss_b200 import features, modules as M
    W = WORKLOAD
    ev = lambda: torch.cuda.Event(enable_timing=True)
    names = ['stft', 'rnn_xproj', 'rnn_recurrent', 'query', 'emb_attn_mask', 'mask_istft']
    acc = {n: [] for n in names}
    lib = d.load_library()
    for _ in range(reps):
        t = {n: 0.0 for n in names}
        with torch.no_grad():
            e0, e1 = ev(), ev(); e0.record()
            batch = sep.features(wav)
            e1.record(); torch.cuda.synchronize(); t['stft'] += e0.elapsed_time(e1)
            # encoder, split into its two kernel families
            packed = sep.mix._packed
            rnn = packed.rnn
            x = batch['mix_feas']
            B, T, _ = x.shape
            G, H = 4 if W['cell'] == 'lstm' else 3, W['H']
            from dl4ss_b200 import _lib
            cell = _lib.CELL_LSTM if W['cell'] == 'lstm' else _lib.CELL_GRU
            tc_rec = M.use_tc_recurrence(H, cell)
            ws = M.recurrent_workspace(B, T, H, cell, tc_rec, x.device)
            xproj = torch.empty(B * T, 2 * G * H, device=x.device)
            inp = x
            layers = packed.get()
            fuse = tc_rec and M.use_tensor_cores()
            planes = hmean = None
            Kpy = (2 * H + 63) // 64 * 64
            for li, lw in enumerate(layers):       # the same sequence rnn_forward issues, with events between the halves
                e0, e1, e2 = ev(), ev(), ev(); e0.record()
                if M.use_tensor_cores():
                    a_pl = planes if planes is not None else M.split_bf16(inp.view(B * T, -1))
                    M.linear_tc(a_pl, M.weight_planes(lw['wih']), lw['bias'], B * T, 2 * G * H, inp.shape[-1], out=xproj)
                else:
                    M.linear_fwd(inp.view(B * T, -1), lw['wih'], lw['bias'], 'none', out=xproj)
                e1.record()
                planes = None
                last = li == len(layers) - 1
                if fuse:
                    planes = torch.empty(2, B * T, Kpy, device=x.device, dtype=torch.bfloat16)
                    if Kpy > 2 * H and not M._RNN_YX:
                        planes[:, :, 2 * H:].zero_()
                    if last:
                        hmean = torch.empty(B, 2 * H, device=x.device)
                # as Separator.masks runs it: with the fused head the layer output exists as bf16 planes only (no fp32 copy)
                y = M.recurrent_layer(lw, cell, xproj, B, T, H, ws, tc_rec, None, None, None, planes, hmean if last else None,
                                      want_y=not fuse)
                if y is None:
                    y = M.HiddenStub((B, T, 2 * H), x.device)
                e2.record(); torch.cuda.synchronize()
                t['rnn_xproj'] += e0.elapsed_time(e1); t['rnn_recurrent'] += e1.elapsed_time(e2)
                inp = y
            e0, e1 = ev(), ev(); e0.record()
            q, _ = sep.queries(inp, idx, hmean)
            e1.record(); torch.cuda.synchronize(); t['query'] += e0.elapsed_time(e1)
            lin = sep.mix.Linear
            e0, e1 = ev(), ev(); e0.record()
            masks = M.emb_attn_mask(inp, lin.weight, lin.bias, q, x.shape[2], W['E'], h_planes=planes)
            e1.record(); torch.cuda.synchronize(); t['emb_attn_mask'] += e0.elapsed_time(e1)
            e0, e1 = ev(), ev(); e0.record()
            features.mask_istft(masks, batch['mix_mag'], W['hop'])
            e1.record(); torch.cuda.synchronize(); t['mask_istft'] += e0.elapsed_time(e1)
        for n in names:
            acc[n].append(t[n])
    return {n: float(np.median(v)) for n, v in acc.items()}


def recurrent_concurrent_ms(sep, B, T, inflight, reps=3):
    """Device time of the recurrent launches (all layers) of `inflight` batches running side by side on as many streams,
    as they do in the timed region: CUDA events from the fork to the join.  Returns ms for the whole set."""
    from dl4ss_b200 import _lib, modules as M
    W = WORKLOAD
    G, H = 4 if W['cell'] == 'lstm' else 3, W['H']
    cell = _lib.CELL_LSTM if W['cell'] == 'lstm' else _lib.CELL_GRU
    dev = next(sep.mix.parameters()).device
    layers = sep.mix._packed.get()
    Kpy = (2 * H + 63) // 64 * 64
    sets = []
    for _ in range(inflight):
        sets.append({'xproj': torch.randn(B * T, 2 * G * H, device=dev) * 0.1,
                     'ws': M.recurrent_workspace(B, T, H, cell, True, dev),
                     'planes': torch.zeros(2, B * T, Kpy, device=dev, dtype=torch.bfloat16),
                     'stream': torch.cuda.Stream(dev)})
    best = 1e9
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record()
        for q in sets:
            q['stream'].wait_event(e0)
            with torch.cuda.stream(q['stream']):
                for lw in layers:
                    # as the timed region runs it: the layer output leaves as bf16 planes only
                    M.recurrent_layer(lw, cell, q['xproj'], B, T, H, q['ws'], True, None, None, None, q['planes'], None, want_y=False)
        for q in sets:
            cur.wait_stream(q['stream'])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def M_use_tc(W):
    from dl4ss_b200 import _lib, modules as M
    cell = _lib.CELL_LSTM if W['cell'] == 'lstm' else _lib.CELL_GRU
    return bool(M.use_tc_recurrence(W['H'], cell))


def algorithmic(B):
    """Algorithmic bytes / flops per step (SURVEY 8d figures x utterances per launch)."""
    W = WORKLOAD
    L, S, hop = W['L'], W['S'], W['hop']
    T, F, H, E = 1 + L // hop, W['n_fft'] // 2 + 1, W['H'], W['E']
    G = 4 if W['cell'] == 'lstm' else 3
    stft_bytes = B * (4 * L + 4 * T * F + 8 * T * F)                       # wav in, |X| and complex X out
    istft_bytes = B * (4 * S * T * F + 8 * T * F + 4 * S * hop * (T - 1))  # masks + mixture in, wavs out
    xproj_flops = sum(2.0 * B * T * (F if l == 0 else 2 * H) * 2 * G * H for l in range(W['layers']))
    rec_flops = 2.0 * B * T * H * G * H * 2 * W['layers']
    lin_flops = 2.0 * B * T * 2 * H * F * E
    return {'stft_bytes': stft_bytes, 'istft_bytes': istft_bytes, 'xproj_flops': xproj_flops,
            'rec_flops': rec_flops, 'emb_flops': lin_flops + 2.0 * B * S * T * F * E}


def cpu_oracle_setup(n_utt, seed=1):
    from oracle import modules_ref as mr, synth
    W = WORKLOAD
    T = 1 + W['L'] // W['hop']
    F = W['n_fft'] // 2 + 1
    torch.manual_seed(1)
    rc = mr.RefConfig(HIDDEN_UNITS=W['H'], EMBEDDING_SIZE=W['E'], NUM_LAYERS=W['layers'],
                      is_ComlexMask=W['complex_mask'])
    mods = (mr.MIX_SPEECH(rc, F, T, W['cell'], W['layers']), mr.SPEECH_EMBEDDING(rc, W['num_spk'], W['E'], 2),
            mr.ATTENTION(rc, W['E'], 'dot'), mr.ADDJUST(rc, 2 * W['H'], W['E']))
    rng = np.random.RandomState(seed)
    wav = rng.standard_normal((n_utt, W['L']))
    idx = np.sort(rng.choice(W['num_spk'], (n_utt, W['S'])), axis=1)
    return rc, mods, wav, idx


def cpu_oracle_step(rc, mods, wav, idx):
    """The reference's path on the CPU: per-utterance librosa-semantics STFTs (incl. the reference's
    redundant second mixture STFT, TDAA_beta/predata_fromList.py:194-200), torch-CPU modules with
    the S-fold expand+baddbmm attention, per-source numpy iSTFT."""
    from oracle import modules_ref as mr, stft_ref as sr
    W = WORKLOAD
    feas, phase = [], []
    for w in wav:
        feas.append(np.transpose(np.abs(sr.stft_ref(w, W['n_fft'], W['hop']))))
        phase.append(np.transpose(sr.stft_ref(w, W['n_fft'], W['hop'])))
    feas = torch.from_numpy(np.array(feas, dtype=np.float32))
    with torch.no_grad():
        r = mr.forward_ref(rc, mods[0], mods[1], mods[2], mods[3], feas, idx)
    return mr.reconstruct_ref(r, np.array(phase), W['hop'])


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_utt = args.ref_utts
    rc, mods, wav, idx = cpu_oracle_setup(n_utt)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_oracle_step(rc, mods, wav, idx)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(rc, mods, wav, idx)
    dt = time.perf_counter() - t0
    audio_s = n_utt * WORKLOAD['L'] / SR * args.steps
    v = audio_s / dt
    line = {'impl': 'reference', 'metric': 'separated_audio_seconds_per_second', 'value': v, 'unit': 'audio-s/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args.batch), reference_sample_utterances_per_step=n_utt,
                           launch='CPU: python loop over utterances for STFT/iSTFT (as the reference), torch-CPU modules on the %d-utterance sample' % n_utt,
                           l2='n/a (CPU)'),
            'cpu_baseline': {'value': v, 'unit': 'audio-s/s', 'cores': cores, 'cpu': cpu_model_name(), 'kind': 'port',
                             'sample': '%d utterances x 5 s per step (bounded sample of the batch-256 workload: each step is %d of the 256 utterances)' % (n_utt, n_utt)},
            'e2e': {'value': v, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)
    return 0


def run_train(args):
    """BASELINE configs[3]: one training step per 'step' (feature STFTs of mixture and clean sources, forward
    with saved gates, MSE loss, hand-written backward, NCCL all-reduce of the flat gradient bucket, Adam),
    utterances sharded by batch over the ranks (weak scaling: --train-batch per GPU)."""
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    else:
        torch.cuda.set_device(0)
    device = torch.device('cuda', local_rank if world > 1 else 0)
    import dl4ss_b200 as d
    d.load_library()
    W, B = WORKLOAD, args.train_batch
    sep = build_model(device)                       # same seed on every rank: identical replicas
    step = d.TrainStep(sep.mix, sep.emb, sep.att, sep.adj)
    opt = torch.optim.Adam([{'params': step.parameters()}], lr=2e-4)       # EvalVer.py:537-544
    g = torch.Generator(device=device).manual_seed(11 + rank)
    src = torch.randn(B, W['S'], W['L'], device=device, generator=g)
    src = src / src.abs().amax(2, keepdim=True)
    wav = src.sum(1).contiguous()
    gi = torch.Generator().manual_seed(7 + rank)
    idx = torch.sort(torch.stack([torch.randperm(W['num_spk'], generator=gi)[:W['S']] for _ in range(B)]), 1)[0].to(device)

    def one_step():
        batch = d.prepare_batch(wav, W['n_fft'], W['hop'], False, sources=src)          # K1: mixture + target STFTs
        return step.step(opt, batch['mix_feas'], idx, batch['multi_spk_fea'].contiguous(), global_batch=B * world)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank if world > 1 else 0)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 1)):
        one_step()
    if rank == 0:
        sampler.wait_first()
    barrier()
    n0 = d.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    for _ in range(args.steps):
        loss = one_step()[0]
    e1.record()
    barrier()
    sampler.end()
    ms = e0.elapsed_time(e1)
    launches = d.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    nparams = sum(p.numel() for p in step.parameters())
    if rank == 0:
        audio_s = world * B * W['L'] / SR * args.steps
        cfg = workload_config(B)
        cfg['workload'] = 'TDAA_beta 2-speaker training step (BASELINE configs[3]): STFT -> BLSTM attention masks -> MSE loss -> backward -> gradient all-reduce -> Adam'
        cfg['allreduce_bytes_per_step'] = nparams * 4 if world > 1 else 0
        cfg['l2'] = 'working set >> 126 MB L2 (saved gates 192 MB/layer at 64 utterances)'
        cfg['launch'] = ('eager launches; the BPTT chain of each layer is one persistent kernel (dl4ss_rnn_layer_bwd_tc), '
                         'the forward recurrence one persistent kernel per layer (dl4ss_rnn_layer_tc_fwd)')
        emit({'metric': 'training_audio_seconds_per_second', 'value': audio_s / (ms * 1e-3), 'unit': 'audio-s/s',
                          'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
                          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                          'data': 'synthetic', 'config': cfg, 'gpu_launches': launches, 'clocks': clocks,
                          'loss': float(loss), 'parameters': nparams})
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def training_extra(device, B=256, steps=5, warmup=3, dist=None, rank=0, world=1, brief=False):
    """BASELINE configs[3] next to the headline line, at every N: the training step of run_train() (STFTs -> forward
    with saved gates -> loss -> backward -> gradient all-reduce -> Adam) at B utterances PER GPU (weak scaling), CUDA-event
    timed, max over ranks.  At N > 1 the NCCL all-reduce of the gradient bucket is inside the timed step, launched per
    layer segment underneath the rest of the backward pass; the same step is also timed with one serial all-reduce after
    the backward pass and with no collective at all, so the cost of the exchange and what the overlap hides are visible."""
    import dl4ss_b200 as d
    W = WORKLOAD
    sep = build_model(device)                       # same seed on every rank: identical replicas
    step = d.TrainStep(sep.mix, sep.emb, sep.att, sep.adj)
    opt = torch.optim.Adam([{'params': step.parameters()}], lr=2e-4)       # EvalVer.py:537-544
    g = torch.Generator(device=device).manual_seed(11 + rank)
    src = torch.randn(B, W['S'], W['L'], device=device, generator=g)
    src = src / src.abs().amax(2, keepdim=True)
    wav = src.sum(1).contiguous()
    gi = torch.Generator().manual_seed(7 + rank)
    idx = torch.sort(torch.stack([torch.randperm(W['num_spk'], generator=gi)[:W['S']] for _ in range(B)]), 1)[0].to(device)

    def one_step(mode):
        batch = d.prepare_batch(wav, W['n_fft'], W['hop'], False, sources=src)
        return step.step(opt, batch['mix_feas'], idx, batch['multi_spk_fea'].contiguous(), global_batch=B * world,
                         overlap=(mode == 'overlap'), reduce=(mode != 'none'))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(mode):
        for _ in range(warmup):
            one_step(mode)
        barrier()
        n0 = d.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = one_step(mode)[0]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, (d.launch_count() - n0) // steps, float(loss)

    ms, launches, loss = timed('overlap')
    out = {'workload': 'training step (BASELINE configs[3]): STFT -> BLSTM attention masks -> MSE loss -> backward -> gradient all-reduce -> Adam',
           'batch_per_gpu': B, 'n_gpus': world, 'scaling': 'weak', 'ms_per_step': ms,
           'value': world * B * W['L'] / SR / (ms * 1e-3), 'unit': 'audio-s/s', 'steps': steps, 'warmup': warmup,
           'gpu_launches_per_step': launches, 'loss': loss,
           'allreduce_bytes_per_step': step.reduced_bytes, 'allreduce_segments': len(step.bucket().ranges),
           'collective': 'NCCL all-reduce (sum) of the persistent flat fp32 gradient bucket, one call per segment (head, then '
                         'each recurrent layer top-down) launched as soon as the segment is written' if world > 1 else 'none (1 GPU)'}
    if world > 1 and not brief:
        out['ms_per_step_serial_collective'] = timed('serial')[0]
        out['ms_per_step_no_collective'] = timed('none')[0]
    del step, opt, sep, src, wav
    torch.cuda.empty_cache()
    if not brief:
        # round 1 quoted the step at 64 utterances per GPU (one BPTT chunk); kept next to the 256-utterance figure
        small = training_extra(device, 64, 3, warmup, dist, rank, world, brief=True)
        out['at_64_utterances_per_gpu'] = {'ms_per_step': small['ms_per_step'], 'value': small['value'], 'unit': 'audio-s/s'}
        out['batch_note'] = ('256 utterances per GPU, the inference batch: the forward recurrence serves them in one launch per layer, '
                             'the BPTT kernel walks four 16-utterance tiles per CTA in one launch per layer')
    return out


def library_baseline(device, B, reps=3):
    """The reference's OWN GPU lowering of the model stages, timed on this GPU ("the bar is the library-call
    path"): torch.nn.LSTM (cuDNN persistent RNN) 4x300 bidirectional + nn.Linear + tanh writing the [B,T,F,E] tensor,
    expand().contiguous() S-fold copy, baddbmm dot attention + sigmoid, mask x mixture
    (TDAA_beta/main_run_sstune_EvalVer.py:282-303,216-226,453-470), fp32 with TF32 off.  The reference runs its STFT /
    iSTFT on the CPU, so this comparator covers features -> masks -> predicted spectra only; it is compared with the sum
    of our encoder + query + fused attention stages.  A comparator, not part of the product: plain torch, nothing of ours."""
    W = WORKLOAD
    T, F, H, E, S = 1 + W['L'] // W['hop'], W['n_fft'] // 2 + 1, W['H'], W['E'], W['S']
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(1)
        rnn = torch.nn.LSTM(F, H, W['layers'], batch_first=True, bidirectional=True).to(device)
        lin = torch.nn.Linear(2 * H, F * E).to(device)
        table = torch.nn.Embedding(W['num_spk'], E).to(device)
        adj = torch.nn.Linear(2 * H + E, E, bias=False).to(device)
        feas = torch.rand(B, T, F, device=device) * 3
        idx = torch.randint(0, W['num_spk'], (B, S), device=device)

        def step():
            x, _ = rnn(feas)
            x = x.contiguous()
            emb = torch.tanh(lin(x.view(B * T, -1))).view(B, T, F, E)
            q = table(idx)
            hm = torch.mean(x, 1).view(B, 1, 2 * H).expand(B, S, 2 * H)
            q = adj(torch.cat([hm, q], 2)) + q
            h5 = emb.view(B, 1, T, F, E).expand(B, S, T, F, E).contiguous().view(-1, T * F, E)
            energy = torch.baddbmm(torch.zeros(B * S, T * F, 1, device=device), h5, q.view(B * S, E, 1))
            mask = torch.sigmoid(energy).view(B, S, T, F)
            return mask * feas.view(B, 1, T, F)

        with torch.no_grad():
            step()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
        return {'ms_per_step': best, 'value': B * W['L'] / SR / (best * 1e-3), 'unit': 'audio-s/s', 'batch': B,
                'what': 'torch.nn.LSTM (cuDNN) 4x300 + Linear + tanh + expand/contiguous + baddbmm + sigmoid + mask*mix, fp32 '
                        '(TF32 off), features -> predicted spectra only (the reference runs STFT/iSTFT on the CPU)',
                'torch': torch.__version__, 'cudnn': torch.backends.cudnn.version()}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()


def cpu_model_name():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.lower().startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def cpu_batched_baseline(n_utt, budget_s=12.0):
    """A CPU baseline that is NOT a strawman: the same model in plain torch on all host cores, but batched the way a
    careful CPU implementation would be -- ONE torch.stft over the batch (no per-utterance python loop, no redundant second
    and third mixture STFT), the modules on the whole batch, one batched torch.istft.  Bounded sample of the workload."""
    W = WORKLOAD
    T, F, H, E, S = 1 + W['L'] // W['hop'], W['n_fft'] // 2 + 1, W['H'], W['E'], W['S']
    torch.manual_seed(1)
    rnn = torch.nn.LSTM(F, H, W['layers'], batch_first=True, bidirectional=True)
    lin = torch.nn.Linear(2 * H, F * E)
    table = torch.nn.Embedding(W['num_spk'], E)
    adj = torch.nn.Linear(2 * H + E, E, bias=False)
    wav = torch.randn(n_utt, W['L'])
    idx = torch.randint(0, W['num_spk'], (n_utt, S))
    win = torch.hann_window(W['n_fft'], periodic=True)

    def step():
        X = torch.stft(wav, W['n_fft'], W['hop'], W['n_fft'], win, center=True, pad_mode='reflect', return_complex=True)
        feas = X.abs().transpose(1, 2).contiguous()                                     # [B,T,F]
        x, _ = rnn(feas)
        q = table(idx)
        q = adj(torch.cat([x.mean(1, keepdim=True).expand(n_utt, S, 2 * H), q], 2)) + q
        emb = torch.tanh(lin(x.reshape(n_utt * T, -1))).view(n_utt, T * F, E)
        mask = torch.sigmoid(torch.bmm(emb, q.transpose(1, 2))).view(n_utt, T, F, S)      # no S-fold copy of the embedding
        Y = mask.permute(0, 3, 2, 1) * X.unsqueeze(1)                                   # [B,S,F,T]
        return torch.istft(Y.reshape(n_utt * S, F, T), W['n_fft'], W['hop'], W['n_fft'], win, center=True,
                           length=W['hop'] * (T - 1))

    with torch.no_grad():
        step()
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t0 < budget_s and reps < 50):
            step()
            reps += 1
        dt = time.perf_counter() - t0
    return {'value': n_utt * W['L'] / SR * reps / dt, 'unit': 'audio-s/s', 'cores': torch.get_num_threads(), 'kind': 'port-batched',
            'sample': '%d reps of one %d-utterance batch x 5 s: batched torch.stft / modules / torch.istft, no redundant STFTs' % (reps, n_utt)}


INFLIGHT = 2          # steps in flight (set from --inflight)


def workload_config(B):
    W = WORKLOAD
    return {'workload': W['name'], 'batch_per_gpu': B, 'utterance_s': W['L'] / SR, 'sample_rate': SR,
            'n_fft': W['n_fft'], 'hop': W['hop'], 'speakers': W['S'],
            'encoder': '%s %dx%d bidirectional' % (W['cell'].upper(), W['layers'], W['H']),
            'embedding': W['E'], 'attention': 'dot + ADDJUST self-tune', 'mask': 'real sigmoid',
            'l2': 'working set >> 126 MB L2 (xproj 769 MB/layer at B=256) and 4 rotating input batches',
            'launch': 'one CUDA graph per step; %d steps (256-utterance batches) in flight on as many streams (PipelinedSeparator / HostPipeline: the recurrent launches take 60 SMs per batch, so two batches share the GPU); the rotating batch is copied device-to-device into its static input inside the timed region' % INFLIGHT}


RESULT = None     # the process's real stdout, reserved for the one JSON line


def emit(line):
    out = RESULT if RESULT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def reserve_stdout():
    """Libraries print to stdout behind our back (NCCL's version banner at communicator creation): keep the real stdout
    for the JSON line and point fd 1 at stderr for everything else."""
    global RESULT
    if RESULT is None:
        sys.stdout.flush()
        RESULT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='utterances per GPU per step')
    ap.add_argument('--ref-utts', type=int, default=32, help='utterances per step of the CPU reference arm')
    ap.add_argument('--cpu-baseline-utts', type=int, default=4)
    ap.add_argument('--cpu-batched-utts', type=int, default=32)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--graph', type=int, default=1, help='1: replay the step from one CUDA graph (default); 0: eager launches')
    ap.add_argument('--inflight', type=int, default=2,
                    help='steps (batches) in flight on as many streams in the inference timings (1: one batch owns the GPU)')
    ap.add_argument('--e2e-depth', type=int, default=0, help='device slots of the end-to-end HostPipeline (0: inflight + 1)')
    ap.add_argument('--no-train-extra', action='store_true', help='skip the configs[3] training-step timing added to the line at N=1')
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'],
                    help="'train': BASELINE configs[3], STFT -> encoder -> masks -> loss -> backward -> all-reduce -> Adam")
    ap.add_argument('--train-batch', type=int, default=256, help='utterances per GPU per training step')
    args = ap.parse_args()
    global INFLIGHT
    INFLIGHT = max(1, args.inflight)
    args.inflight = INFLIGHT
    if args.impl == 'reference':
        return run_reference(args)
    if args.mode == 'train':
        return run_train(args)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    else:
        torch.cuda.set_device(0)
    device = torch.device('cuda', local_rank if world > 1 else 0)
    import dl4ss_b200 as d
    d.load_library()          # raises if the CUDA library is missing: no fallback

    B, W = args.batch, WORKLOAD
    sep = build_model(device)
    wavs = make_wave_batches(B, W['L'], 4, device, seed=1 + rank)
    g = torch.Generator().manual_seed(7 + rank)
    idx = torch.sort(torch.stack([torch.randperm(W['num_spk'], generator=g)[:W['S']] for _ in range(B)]), 1)[0].to(device)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): the step is one CUDA graph (GraphedSeparator), replayed per batch;
    # the rotating input batch is copied device-to-device into the graph's static input inside the timed region
    sampler = ClockSampler(local_rank if world > 1 else 0)
    if rank == 0:
        sampler.start()
    sep.separate(wavs[0], idx, check_index=False)      # first use: weight planes, packed W_hh, twiddles
    n0 = d.launch_count()
    sep.separate(wavs[1 % len(wavs)], idx, check_index=False)
    launches_per_step = d.launch_count() - n0          # kernels of ours in one step (the graph replays exactly these)
    if args.graph:
        # `--inflight` 256-utterance batches on as many streams, one CUDA graph each (PipelinedSeparator); a step's result is
        # picked up on the timing stream one submit later, as a consumer of the separated waveforms would
        pipe_dev = d.PipelinedSeparator(sep, B, W['L'], W['S'], depth=args.inflight, device=device)
        pending = []

        def step_fn(w, i):
            pending.append(pipe_dev.submit(w, i))
            while len(pending) >= args.inflight:
                k = pending.pop(0)
                pipe_dev.result(k)
                pipe_dev.release(k)

        def finish():
            while pending:
                k = pending.pop(0)
                pipe_dev.result(k)
                pipe_dev.release(k)
    else:
        step_fn = lambda w, i: sep.separate(w, i, check_index=False)
        finish = lambda: None
    for i in range(args.warmup):
        step_fn(wavs[i % len(wavs)], idx)
    finish()
    if rank == 0:
        sampler.wait_first()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    for i in range(args.steps):
        step_fn(wavs[i % len(wavs)], idx)
    finish()
    e1.record()
    barrier()
    sampler.end()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public call with HOST buffers ("e2e"): pinned waveforms in, pinned separated
    # waveforms out, every step; HostPipeline overlaps the copies of neighbouring steps with the kernels
    h_in = [w.cpu().pin_memory() for w in wavs]
    h_idx = idx.cpu().pin_memory()
    Lout = W['hop'] * (W['L'] // W['hop'])
    h_outs = [torch.empty(B, W['S'], Lout, dtype=torch.float32).pin_memory() for _ in range((args.e2e_depth or args.inflight + 1) + 1)]
    pipe = d.HostPipeline(sep, B, W['L'], W['S'], depth=args.e2e_depth or args.inflight + 1, device=device, graphs=bool(args.graph),
                          concurrent=args.inflight > 1)
    for i in range(max(8, args.warmup)):      # also lets the PCIe link leave its idle (down-trained) state
        pipe.submit(h_in[i % 4], h_idx, h_outs[i % len(h_outs)])
    pipe.drain()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        pipe.submit(h_in[i % 4], h_idx, h_outs[i % len(h_outs)])
    pipe.drain()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    e2e_check = float(h_outs[(args.steps - 1) % len(h_outs)].abs().max())      # the result really is on the host
    assert e2e_check > 0.0 and e2e_check == e2e_check

    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device=device, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])

    audio_s = world * B * W['L'] / SR * args.steps
    value = audio_s / (ms * 1e-3)
    e2e_v = audio_s / (ms_e2e * 1e-3)

    train = None
    if not args.no_train_extra:          # every rank takes part: at N > 1 the NCCL gradient all-reduce is inside the timed step
        try:
            train = training_extra(device, dist=dist, rank=rank, world=world)
        except Exception as e:           # the headline line must not depend on the extra measurement
            train = {'error': repr(e)[:200]}

    line = None
    if rank == 0:
        peaks = measured_peaks()
        # stage times: one launch at a time with the whole GPU to itself (what the kernels themselves achieve); the timed
        # region runs `inflight` batches side by side with the shared-SM geometry, measured for the dominant stage below
        st = stage_times(sep, wavs[0], idx)
        rec_conc = None
        if args.graph and args.inflight > 1 and M_use_tc(W):
            with d.sm_sharing(*d.sm_sharing.PIPELINED):
                rec_conc = recurrent_concurrent_ms(sep, B, 1 + W['L'] // W['hop'], 2)
        st.update(kernel_stage_times(sep, wavs, idx))
        alg = algorithmic(B)
        flops = {'rnn_xproj': alg['xproj_flops'], 'rnn_recurrent': alg['rec_flops'], 'emb_attn_mask': alg['emb_flops']}
        dom = max(flops, key=lambda k: st[k])
        n_launch = {'rnn_xproj': W['layers'], 'rnn_recurrent': W['layers'], 'emb_attn_mask': 1}[dom]
        ach = flops[dom] / (st[dom] * 1e-3) / 1e12
        conc = None
        if dom == 'rnn_recurrent' and rec_conc is not None:
            # two batches' recurrent launches run side by side in the timed region (60 SMs each): the GPU-level figure is
            # their joint work over their joint duration; the figure of one launch owning the GPU (80 SMs) is kept next to it
            conc = {'batches': 2, 'ctas_per_launch': 60, 'ms_all_layers': rec_conc,
                    'single_launch_achieved': ach, 'single_launch_frac': ach / peaks['bf16_tflops_sustained'],
                    # strict per-launch reading: one batch's launches while the other batch's run beside them on the other SMs
                    'per_launch_in_region_achieved': flops[dom] / (rec_conc * 1e-3) / 1e12,
                    'per_launch_in_region_frac': flops[dom] / (rec_conc * 1e-3) / 1e12 / peaks['bf16_tflops_sustained'],
                    'explanation': 'achieved / frac above are GPU level: two batches\' recurrent launches (60 CTAs each) run side by side in the '
                                   'timed region, so the GPU completes 2 x flops per ms_all_layers; single_launch_* is one launch owning the GPU '
                                   '(80 CTAs), per_launch_in_region_* one of the two concurrent launches taken alone'}
            ach = 2 * flops[dom] / (rec_conc * 1e-3) / 1e12
        tr = ncu_traffic({'rnn_recurrent': 'rnn_tc_kernel', 'rnn_xproj': 'EpiPlain', 'emb_attn_mask': 'EpiAttn'}[dom])
        roof = {'kernel': dom, 'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_tflops_sustained'],
                'unit': 'TFLOP/s', 'frac': ach / peaks['bf16_tflops_sustained'],
                'traffic': tr['bytes_per_launch'] if tr else None, 'traffic_source': tr['source'] if tr else None,
                # K3 per launch as the product path runs it: reads xproj [B,T,2,G*H] fp32, writes the layer output as bf16 hi/lo
                # planes [2][B*T][2H] (no fp32 copy; h is exchanged through these rows)
                'algorithmic_bytes_per_launch': (B * (1 + W['L'] // W['hop']) * (2 * (4 if W['cell'] == 'lstm' else 3) * W['H'] * 4
                                                 + 2 * (2 * W['H']) * 2)) if dom == 'rnn_recurrent' else None,
                'peak_source': peaks['source'] + ' (sustained cuBLAS bf16; kernel timed inside a long step)',
                'launches_per_step': n_launch, 'ms_per_step': st[dom], 'concurrent': conc,
                'note': 'algorithmic fp32 FLOPs; every tensor-core stage runs bf16x3 (3 MMA products per fp32 product); the recurrent stage is a latency chain of T sequential steps per layer, not throughput bound (DESIGN.md 4)'}
        stages = {}
        for k, by in (('stft', alg['stft_bytes']), ('mask_istft', alg['istft_bytes'])):
            a = by / (st[k] * 1e-3) / 1e9
            stages[k] = {'bound': 'hbm', 'achieved': a, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                         'frac': a / peaks['hbm_gbs'], 'ms': st[k], 'bytes': by}
        for k in ('rnn_xproj', 'rnn_recurrent', 'emb_attn_mask'):
            a = flops[k] / (st[k] * 1e-3) / 1e12
            stages[k] = {'bound': 'tensor', 'achieved': a, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                         'frac': a / peaks['bf16_tflops_sustained'], 'ms': st[k], 'flops': flops[k]}
        stages['query'] = {'ms': st['query']}
        cpu = cpu_batched = lib_base = None
        if world == 1:
            try:
                lib_base = library_baseline(device, B)
                ours_model_ms = st['rnn_xproj'] + st['rnn_recurrent'] + st['query'] + st['emb_attn_mask']
                lib_base['ours_same_stages_ms'] = ours_model_ms
                lib_base['speedup_vs_library'] = lib_base['ms_per_step'] / ours_model_ms
            except Exception as e:
                lib_base = {'error': repr(e)[:200]}
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (the other ranks would spin in a barrier meanwhile)
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            n_utt = args.cpu_baseline_utts
            rc, mods, wav, cidx = cpu_oracle_setup(n_utt)
            cpu_oracle_step(rc, mods, wav[:1], cidx[:1])
            t0 = time.perf_counter()
            reps = 0
            while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 50):
                cpu_oracle_step(rc, mods, wav, cidx)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {'value': n_utt * W['L'] / SR * reps / dt, 'unit': 'audio-s/s', 'cores': cores, 'cpu': cpu_model_name(), 'kind': 'port',
                   'sample': '%d reps of %d utterances x 5 s (same model/config, bounded sample)' % (reps, n_utt)}
            try:
                cpu_batched = cpu_batched_baseline(args.cpu_batched_utts)
                cpu_batched['cpu'] = cpu['cpu']
            except Exception as e:
                cpu_batched = {'error': repr(e)[:200]}
        bytes_in = B * W['L'] * 4 + idx.numel() * 8
        bytes_out = B * W['S'] * Lout * 4
        line = {'metric': 'separated_audio_seconds_per_second', 'value': value, 'unit': 'audio-s/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': workload_config(B),
                'e2e': {'value': e2e_v, 'unit': 'audio-s/s', 'h2d_bytes_per_step': bytes_in,
                        'd2h_bytes_per_step': bytes_out, 'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'roofline_stages': stages,
                'cpu_baseline': cpu, 'cpu_baseline_batched': cpu_batched, 'library_baseline': lib_base,
                'configs3_training': train}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

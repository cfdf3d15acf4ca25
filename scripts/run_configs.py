"""Time every BASELINE.json config through the public pipeline on one B200 (device-resident inputs, CUDA events,
3 warm-up + 10 timed calls over 4 rotating batches) and print a markdown table.

    python scripts/run_configs.py [--out profiles/configs_TAG.md]
configs[0] Torch_multi 2-speaker forward: 8 kHz, hop 64, BLSTM 2x300, batch 8 x 4 s
configs[1] TDAA_beta 2-speaker, LSTM 4x300, batch 256 x 5 s            (the bench.py workload)
configs[2] 3-speaker cRM complex masks, GRU 2x300, batch 512 x 5 s
configs[3] training step (see bench.py --mode train)
configs[4] STFT/iSTFT sweep (see scripts/sweep_stft.py)
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dl4ss_b200 as d

ap = argparse.ArgumentParser()
ap.add_argument('--out', default=None)
args = ap.parse_args()
dev = torch.device('cuda:0')
CFG = [('configs[0] Torch_multi 2-spk, hop 64, LSTM 2x300', dict(B=8, L=32000, S=2, hop=64, cell='lstm', layers=2, cplx=False)),
       ('configs[0] at hop 128 (the reference value)', dict(B=8, L=32000, S=2, hop=128, cell='lstm', layers=2, cplx=False)),
       ('configs[1] TDAA_beta 2-spk, LSTM 4x300', dict(B=256, L=40000, S=2, hop=128, cell='lstm', layers=4, cplx=False)),
       ('configs[2] 3-spk cRM, GRU 2x300', dict(B=512, L=40000, S=3, hop=128, cell='gru', layers=2, cplx=True))]
rows = []
for name, c in CFG:
    torch.manual_seed(1)
    d.config.HIDDEN_UNITS, d.config.EMBEDDING_SIZE = 300, 50
    d.config.is_ComlexMask, d.config.is_SelfTune, d.config.FRAME_SHIFT = c['cplx'], True, c['hop']
    T = 1 + c['L'] // c['hop']
    mix = d.MIX_SPEECH(129, T, cell=c['cell'], num_layers=c['layers']).to(dev)
    emb = d.SPEECH_EMBEDDING(101, 50, c['S']).to(dev)
    if c['cplx']:
        with torch.no_grad():
            emb.layer.weight.mul_(0.2)          # keep the cRM energies in the range where the decompression is finite
    att = d.ATTENTION(50, 'dot').to(dev)
    adj = d.ADDJUST(600, 50).to(dev)
    sep = d.Separator(mix, emb, att, adj, 256, c['hop'])
    g = torch.Generator(device=dev).manual_seed(3)
    wavs = [torch.randn(c['B'], c['L'], device=dev, generator=g) * 0.3 for _ in range(4)]
    idx = torch.sort(torch.stack([torch.randperm(101)[:c['S']] for _ in range(c['B'])]), 1)[0].to(dev)
    for i in range(3):
        out = sep.separate(wavs[i % 4], idx, check_index=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        out = sep.separate(wavs[i % 4], idx, check_index=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    finite = bool(torch.isfinite(out).all())
    # the same calls with two batches in flight (PipelinedSeparator: two streams, one CUDA graph each)
    pipe = d.PipelinedSeparator(sep, c['B'], c['L'], c['S'], depth=2, device=dev)
    pend = []
    def sub(i):
        pend.append(pipe.submit(wavs[i % 4], idx))
        while len(pend) >= 2:
            k = pend.pop(0); pipe.result(k); pipe.release(k)
    for i in range(4):
        sub(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        sub(i)
    while pend:
        k = pend.pop(0); pipe.result(k); pipe.release(k)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 12
    del pipe
    rows.append((name, c, ms, c['B'] * c['L'] / 8000.0 / (ms * 1e-3), finite, ms2))
    print(name, round(ms, 3), 'ms', round(rows[-1][3]), 'audio-s/s', 'finite' if finite else 'NON-FINITE', flush=True)
    d.config.is_ComlexMask, d.config.FRAME_SHIFT = False, 128
lines = ['| config | batch x seconds | speakers | hop | encoder | mask | ms per call (one batch at a time, eager) | audio-s/s | ms per call, two in flight (graphs) | audio-s/s |',
         '|---|---|---:|---:|---|---|---:|---:|---:|---:|']
for name, c, ms, thr, fin, ms2 in rows:
    lines.append('| %s | %d x %.0f | %d | %d | %s %dx300 | %s | %.3f | %.0f | %.3f | %.0f |' % (
        name, c['B'], c['L'] / 8000.0, c['S'], c['hop'], c['cell'].upper(), c['layers'], 'cRM' if c['cplx'] else 'real', ms, thr,
        ms2, c['B'] * c['L'] / 8000.0 / (ms2 * 1e-3)))
txt = '# BASELINE configs through `Separator.separate` on one B200 (device-resident waveforms -> separated waveforms, fp32 results)\n\n' + '\n'.join(lines) + '\n'
print(txt)
if args.out:
    open(args.out, 'w').write(txt)

"""BASELINE configs[4]: STFT -> mask-apply -> iSTFT sweep over utterance length x batch, achieved HBM GB/s.

    python scripts/sweep_stft.py [--out profiles/sweep_stft_TAG.md]

Per point: K1 (waveform -> |X| + complex X) and K6 (S=2 real masks x X -> waveforms), each captured in a CUDA
graph (4 launches over rotating buffers when the working set is below the 126 MB L2) and replayed; algorithmic
bytes per SURVEY 8(d); peak = MEASURED_PEAKS.json hbm_gbs."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dl4ss_b200 as d
from dl4ss_b200 import features

ap = argparse.ArgumentParser()
ap.add_argument('--out', default=None)
ap.add_argument('--hop', type=int, default=128)
ap.add_argument('--S', type=int, default=2)
ap.add_argument('--crm', action='store_true', help='complex masks [B,S,T,F,2] (BASELINE configs[2])')
ap.add_argument('--points', default=None, help='comma list of secs:batch')
args = ap.parse_args()
dev = torch.device('cuda:0')
peak = 6548.8
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p)).get('hbm_gbs', peak)
S, hop = args.S, args.hop
rows = []
points = [(1, 1), (1, 64), (1, 4096), (5, 1), (5, 16), (5, 256), (5, 1024), (5, 4096), (30, 1), (30, 64), (30, 512)]
if args.points:
    points = [tuple(int(v) for v in q.split(':')) for q in args.points.split(',')]
MB = 8 if args.crm else 4
for secs, B in points:
    L = secs * 8000
    T, F = 1 + L // hop, 129
    nbuf = 4 if B * L * 4 * 4 < 2e9 else 2
    wavs = [torch.randn(B, L, device=dev) for _ in range(nbuf)]
    feat = [torch.empty(B, T, F, device=dev) for _ in range(2)]
    cplx = [torch.empty(B, T, F, 2, device=dev) for _ in range(2)]
    masks = torch.rand(B, S, T, F, 2, device=dev) if args.crm else torch.rand(B, S, T, F, device=dev)
    out = [torch.empty(B, S, hop * (T - 1), device=dev) for _ in range(2)]
    specs = []
    for w in wavs:
        _, c = features.stft_features(w, 256, hop, 'hann', None)
        specs.append(c)
    fns = {'stft': lambda i: features.stft_features(wavs[i % nbuf], 256, hop, 'hann', 'abs', out_feat=feat[i % 2], out_cplx=cplx[i % 2]),
           'istft': lambda i: features.mask_istft(masks, specs[i % nbuf], hop, out=out[i % 2])}
    byts = {'stft': B * (4 * L + 12 * T * F), 'istft': B * (MB * S * T * F + 8 * T * F + 4 * S * hop * (T - 1))}
    res = {}
    for name, fn in fns.items():
        fn(0); fn(1)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        inner = 4
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for i in range(inner):
                    fn(i)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / inner)
        res[name] = (best, byts[name] / (best * 1e-3) / 1e9)
    rows.append((secs, B, res))
    print(secs, B, {k: (round(v[0], 4), round(v[1], 1)) for k, v in res.items()}, flush=True)
    del wavs, feat, cplx, masks, out, specs
    torch.cuda.empty_cache()
lines = ['| utterance s | batch | STFT ms | STFT GB/s | frac of %.0f | mask+iSTFT ms | GB/s | frac |' % peak, '|---:|---:|---:|---:|---:|---:|---:|---:|']
for secs, B, r in rows:
    lines.append('| %d | %d | %.4f | %.0f | %.3f | %.4f | %.0f | %.3f |' % (secs, B, r['stft'][0], r['stft'][1], r['stft'][1] / peak,
                                                                        r['istft'][0], r['istft'][1], r['istft'][1] / peak))
txt = '\n'.join(lines)
print(txt)
if args.out:
    open(args.out, 'w').write('# STFT / mask+iSTFT sweep (hop %d, S=%d, %s masks), 1 B200, CUDA-graph replay timing\n\n' % (hop, S, 'complex (cRM)' if args.crm else 'real') + txt + '\n')

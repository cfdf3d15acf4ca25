"""BASELINE configs[4]: STFT -> mask-apply -> iSTFT sweep over utterance length x batch, achieved HBM GB/s.

    python scripts/sweep_stft.py [--hop 64|128] [--S 2|3] [--crm] [--out profiles/sweep_stft_TAG.md] [--append]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep_stft.py ...
        (every rank sweeps its own shard of `batch` utterances per GPU; a point's time is the max over ranks after a
         barrier, the GB/s column is the aggregate of all GPUs, `frac` is per GPU)

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
ap.add_argument('--append', action='store_true', help='append to --out instead of replacing it')
args = ap.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
dist = None
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
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
            if dist is not None:
                dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / inner
            if dist is not None:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            best = min(best, ms)
        res[name] = (best, world * byts[name] / (best * 1e-3) / 1e9)
    rows.append((secs, B, res))
    if rank == 0:
        print(secs, B, {k: (round(v[0], 4), round(v[1], 1)) for k, v in res.items()}, flush=True)
    del wavs, feat, cplx, masks, out, specs
    torch.cuda.empty_cache()
lines = ['| utterance s | batch per GPU | STFT ms | STFT GB/s (all GPUs) | frac of %.0f per GPU | mask+iSTFT ms | GB/s (all GPUs) | frac per GPU |' % peak,
         '|---:|---:|---:|---:|---:|---:|---:|---:|']
for secs, B, r in rows:
    lines.append('| %d | %d | %.4f | %.0f | %.3f | %.4f | %.0f | %.3f |' % (secs, B, r['stft'][0], r['stft'][1], r['stft'][1] / peak / world,
                                                                        r['istft'][0], r['istft'][1], r['istft'][1] / peak / world))
txt = '\n'.join(lines)
if rank == 0:
    print(txt)
    if args.out:
        open(args.out, 'a' if args.append else 'w').write(
            '## hop %d, S=%d, %s masks, %d B200 (CUDA-graph replay timing, max over ranks)\n\n' %
            (hop, S, 'complex (cRM)' if args.crm else 'real', world) + txt + '\n\n')
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()

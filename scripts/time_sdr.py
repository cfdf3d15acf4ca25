"""Time the on-device BSS-Eval (n2) at the BASELINE configs[1] batch: 256 utterances x 2 sources x 39936 samples."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl4ss_b200 import metrics
B, S, N = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 2, 39936
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(1)
ref = torch.randn(B, S, N, device=dev, generator=g)
ref = torch.nn.functional.avg_pool1d(ref, 5, 1, 2)            # coloured
est = ref.flip(1) + 0.3 * ref + 0.1 * torch.randn(B, S, N, device=dev, generator=g)
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    rr = metrics.xcorr_f64(ref, ref, 1023, -511)
    e[1].record()
    rd = metrics.xcorr_f64(ref, est, 512, 0)
    e[2].record()
    ee = (est.double() ** 2).sum(-1)
    out = metrics.bss_from_correlations(rr, rd, ee)
    e[3].record()
    torch.cuda.synchronize()
    print('xcorr ref.ref %.2f ms, ref.est %.2f ms, gram+cholesky+energies %.2f ms, wall %.1f ms' % (
        e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), (time.time() - t) * 1e3), flush=True)
flop = 2.0 * B * S * S * N * (1023 + 512)
print('xcorr: %.1f GFLOP fp64 per batch; mean SDR %.2f dB' % (flop / 1e9, float(out[0].mean())))

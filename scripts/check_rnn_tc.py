"""Quick GPU check of the tcgen05 recurrent kernel: parity vs nn.LSTM/nn.GRU (CPU fp32) and timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dl4ss_b200 as d
from tests.util import build_pair

dev = torch.device('cuda:0')
cases = [('lstm', 1, 3, 5), ('lstm', 2, 3, 37), ('gru', 2, 5, 29), ('lstm', 4, 70, 25), ('gru', 2, 130, 21), ('lstm', 1, 64, 9)]
if len(sys.argv) > 1 and sys.argv[1] == 'one':
    cases = cases[:1]
for cell, layers, B, T in cases:
    ref, ours = build_pair(cell, layers, 129, T, False)
    torch.manual_seed(5)
    x = torch.rand(B, T, 129) * 2
    with torch.no_grad():
        y_ref, _ = ref['mix'].layer(x)
        d.config.RNN_TENSOR_CORES = True
        y = ours['mix'].encode(x.to(dev)).cpu()
        d.config.RNN_TENSOR_CORES = False
        y2 = ours['mix'].encode(x.to(dev)).cpu()
    print(cell, layers, B, T, 'tc err %.3e  simt err %.3e' % ((y - y_ref).abs().max().item(), (y2 - y_ref).abs().max().item()), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == 'one':
    sys.exit(0)
# timing at the bench shape
ref, ours = build_pair('lstm', 4, 129, 313, False)
x = torch.rand(256, 313, 129, device=dev)
for flag in (True, False):
    d.config.RNN_TENSOR_CORES = flag
    with torch.no_grad():
        for _ in range(2):
            ours['mix'].encode(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ours['mix'].encode(x)
        e1.record(); torch.cuda.synchronize()
    print('encode B=256 T=313 LSTM4x300 tc=%s: %.3f ms' % (flag, e0.elapsed_time(e1) / 5), flush=True)

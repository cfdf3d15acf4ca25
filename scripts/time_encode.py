"""Encoder (4 x BLSTM 300) time at several batch sizes / tiles-per-CTA settings of the tcgen05 recurrent kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dl4ss_b200 as d
from dl4ss_b200 import _lib
from tests.util import build_pair

dev = torch.device('cuda:0')
lib = _lib.load()
ref, ours = build_pair('lstm', 4, 129, 313, False)
for B, tpc in [(256, 0), (256, 3), (512, 0), (128, 0), (128, 2), (64, 0), (8, 0)]:
    lib.dl4ss_rnn_tc_set_tiles_per_cta(tpc)
    x = torch.rand(B, 313, 129, device=dev)
    with torch.no_grad():
        for _ in range(2):
            ours['mix'].encode(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ours['mix'].encode(x)
        e1.record(); torch.cuda.synchronize()
    print('encode B=%d tiles/CTA>=%d: %.3f ms' % (B, tpc, e0.elapsed_time(e1) / 5), flush=True)
lib.dl4ss_rnn_tc_set_tiles_per_cta(0)

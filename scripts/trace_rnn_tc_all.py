"""Step-to-step cycles of the tcgen05 recurrent kernel (CTA 0, tile 0) over ALL T steps, in windows of 32 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
import dl4ss_b200 as d
from dl4ss_b200 import _lib
from tests.util import build_pair

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 313
ref, ours = build_pair('lstm', 1, 129, T, False)
x = torch.rand(B, T, 129, device=dev)
lib = _lib.load()
steps = T
buf = torch.zeros(steps * 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    ours['mix'].encode(x)
    torch.cuda.synchronize()
    lib.dl4ss_rnn_tc_set_trace(ctypes.c_void_p(buf.data_ptr()), steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ours['mix'].encode(x); e1.record()
    torch.cuda.synchronize()
    print('encode (K1-less: K2 + K3, 1 layer) %.3f ms' % e0.elapsed_time(e1))
    lib.dl4ss_rnn_tc_set_trace(None, 0)
t = buf.cpu().numpy().reshape(steps, 16)
pd = t[1:, 1]
print('entry -> W resident: %d cycles; W resident -> first poll_done: %d; last poll_done -> all done: %d; entry -> done: %d' % (
    t[0, 3] - t[0, 2], t[1, 1] - t[0, 3], t[0, 4] - t[-1, 1], t[0, 4] - t[0, 2]))
print('total cycles first->last poll_done: %d over %d steps = %.0f per step' % (pd[-1] - pd[0], len(pd) - 1, (pd[-1] - pd[0]) / (len(pd) - 1)))
for w in range(0, len(pd) - 1, 32):
    seg = pd[w:w + 33]
    rel = t[1 + w:1 + w + 32]
    print('steps %3d-%3d: %.0f cycles/step   poll wait %.0f  h landed %.0f  red %.0f' % (
        w + 1, w + len(seg) - 1, np.diff(seg).mean(), (rel[:, 1] - rel[:, 0]).mean(), (rel[:, 4] - rel[:, 1]).mean(),
        (rel[:, 14] - rel[:, 13]).mean()))

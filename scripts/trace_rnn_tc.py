"""Per-phase clock stamps of the tcgen05 recurrent kernel (CTA 0) at the bench shape."""
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
CELL = sys.argv[2] if len(sys.argv) > 2 else 'lstm'
ref, ours = build_pair(CELL, 1, 129, T, False)
x = torch.rand(B, T, 129, device=dev)
lib = _lib.load()
steps = 64
buf = torch.zeros(steps * 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    ours['mix'].encode(x)
    torch.cuda.synchronize()
    lib.dl4ss_rnn_tc_set_trace(ctypes.c_void_p(buf.data_ptr()), steps)
    ours['mix'].encode(x)
    torch.cuda.synchronize()
    lib.dl4ss_rnn_tc_set_trace(None, 0)
t = buf.cpu().numpy().reshape(steps, 16)
names = ['poll_start', 'poll_done', 'tma_issued', 'h0_landed', 'hlast_landed', 'mma_committed', 'tfull_seen', 'tmem_read',
         'transposed', 'math_done', 'stores_done', 'xfull_seen', 'rel_at_barrier', 'rel_barrier_done', 'red_done', 'x_read']
print('step-to-step (poll_done) cycles:', np.diff(t[10:60, 1]).mean())
base = t[10:60, 1:2]
rel = (t[10:60, :16] - base).astype(np.float64)
# epilogue stamps 6.. of step s belong to the same step as loader stamps of step s
for i, n in enumerate(names):
    print('%-14s %9.0f' % (n, rel[:, i].mean()))
# stamps 9..14 at step s precede poll_done of step s+1
nxt = (t[11:61, 1] - t[10:60, 14]).mean()
print('red_done(s) -> poll_done(s+1): %.0f' % nxt)

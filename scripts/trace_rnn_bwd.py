"""Per-phase clock stamps of the persistent BPTT kernel (CTA 0) at the training bench shape (B utterances, T=313)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
from dl4ss_b200 import _lib as L

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
CELL = sys.argv[2] if len(sys.argv) > 2 else 'lstm'
T, H = 313, 300
G = 4 if CELL == 'lstm' else 3
c = L.CELL_LSTM if CELL == 'lstm' else L.CELL_GRU
lib = L.load()
g = torch.Generator(device=dev).manual_seed(1)
dy = torch.randn(B, T, 2 * H, device=dev, generator=g)
whh = torch.randn(2, G * H, H, device=dev, generator=g) / H ** 0.5
gates = torch.rand(B, T, 2, G * H, device=dev, generator=g)
cells = torch.randn(B, T, 2, H, device=dev, generator=g)
y = torch.rand(B, T, 2 * H, device=dev, generator=g)
dgx = torch.empty(B, T, 2, G * H, device=dev)
dgh = torch.empty(B, T, 2, G * H, device=dev) if G == 3 else None
need = lib.dl4ss_rnn_bwd_workspace_bytes(B, T, H, c)
ws = torch.empty(need, device=dev, dtype=torch.uint8)


TC = (sys.argv[3] if len(sys.argv) > 3 else 'tc') == 'tc'
xp = torch.zeros(lib.dl4ss_rnn_bwd_tc_xplanes_bytes(B, T, H, c), device=dev, dtype=torch.uint8)


def run():
    if TC:
        rc = lib.dl4ss_rnn_layer_bwd_tc(c, L.ptr(dy), L.ptr(whh), L.ptr(gates), L.ptr(cells), L.ptr(y), L.ptr(dgx),
                                        L.ptr(dgh), ctypes.c_void_p(xp.data_ptr()), B, T, H,
                                        ctypes.c_void_p(ws.data_ptr()), need, L.stream())
        L.check(rc, 'dl4ss_rnn_layer_bwd_tc')
        return
    rc = lib.dl4ss_rnn_layer_bwd(c, L.ptr(dy), L.ptr(whh), L.ptr(gates), L.ptr(cells), L.ptr(y), L.ptr(dgx), L.ptr(dgh),
                                 B, T, H, ctypes.c_void_p(ws.data_ptr()), need, L.stream())
    L.check(rc, 'dl4ss_rnn_layer_bwd')


run(); run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
print('B=%d %s %s: %.3f ms per layer launch' % (B, CELL, 'tc' if TC else 'fp32', e0.elapsed_time(e1) / 5))
steps = 64
buf = torch.zeros(steps * 8, dtype=torch.int64, device=dev)
lib.dl4ss_rnn_bwd_set_trace(ctypes.c_void_p(buf.data_ptr()), steps)
run()
torch.cuda.synchronize()
lib.dl4ss_rnn_bwd_set_trace(None, 0)
t = buf.cpu().numpy().reshape(steps, 8)
names = ['poll_start', 'counter_seen', 'dg_loaded', 'product_done', 'partials_synced', 'gates_stored', 'cta_synced', 'released']
print('step-to-step cycles:', np.diff(t[10:60, 1]).mean())
rel = (t[10:60] - t[10:60, 0:1]).astype(np.float64)
for i, n in enumerate(names):
    print('%-16s %9.0f' % (n, rel[:, i].mean()))
print('released(s) -> counter_seen(s+1): %.0f' % (t[11:61, 1] - t[10:60, 7]).mean())

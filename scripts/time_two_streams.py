"""Throughput of the whole separation step with TWO 256-utterance batches in flight on two streams (one CUDA graph
each) for several (tiles per recurrent CTA, projection CTA cap) settings, next to one stream and to one 512 batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dl4ss_b200 as d
from dl4ss_b200 import _lib

dev = torch.device('cuda:0')
lib = _lib.load()
W = bench.WORKLOAD
sep = bench.build_model(dev)
B = 256
g = torch.Generator(device='cpu'); g.manual_seed(1)
idx = torch.sort(torch.stack([torch.randperm(W['num_spk'], generator=g)[:W['S']] for _ in range(B)]), 1)[0].to(dev)
wav = torch.randn(B, W['L'], device=dev)
steps = 12


def run(nstreams, tpc, cap, Bx=B):
    lib.dl4ss_rnn_tc_set_tiles_per_cta(tpc)
    lib.dl4ss_gemm_tc_set_max_ctas(cap)
    w = wav if Bx == B else torch.cat([wav] * (Bx // B))
    ix = idx if Bx == B else torch.cat([idx] * (Bx // B))
    gs = [d.GraphedSeparator(sep, Bx, W['L'], W['S'], device=dev) for _ in range(nstreams)]
    for q in gs:
        q.wav.copy_(w); q.idx.copy_(ix)
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            with torch.cuda.stream(streams[i % nstreams]):
                gs[i % nstreams].graph.replay()
        cur = torch.cuda.current_stream()
        for s in streams:
            cur.wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps * (B / Bx)
    print('streams %d  B %d  tiles/CTA>=%d  gemm cap %3d : %.3f ms per 256 utterances' % (nstreams, Bx, tpc, cap, ms), flush=True)


run(1, 0, 0)
run(1, 0, 0, 512)
run(2, 0, 0)
run(2, 3, 0)
run(2, 3, 120)
run(2, 3, 88)
run(2, 2, 0)
run(3, 3, 0)
lib.dl4ss_rnn_tc_set_tiles_per_cta(0)
lib.dl4ss_gemm_tc_set_max_ctas(0)

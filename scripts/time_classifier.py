import sys; sys.path.insert(0,'.')
import torch, dl4ss_b200 as d
d.config.HIDDEN_UNITS=300
for B in (16, 64, 256):
    cls = d.MIX_SPEECH_classifier(129, 313, 101).cuda()
    x = torch.rand(B, 313, 129, device='cuda')
    with torch.no_grad():
        cls(x); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): p = cls(x)
        e1.record(); torch.cuda.synchronize()
    print('classifier B=%d T=313: %.1f ms' % (B, e0.elapsed_time(e1)/3))

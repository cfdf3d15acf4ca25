"""One classifier forward (BLSTM 3x600) at B utterances: target for ncu captures of rnn_mma_kernel."""
import sys; sys.path.insert(0, '.')
import torch, dl4ss_b200 as d
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cls = d.MIX_SPEECH_classifier(129, 313, 101).cuda()
x = torch.rand(B, 313, 129, device='cuda')
with torch.no_grad():
    for _ in range(2):
        cls(x)
torch.cuda.synchronize()

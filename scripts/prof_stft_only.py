"""ncu target: a few K1 / K6 launches at a fixed size (default 5 s x 1024 utterances, S=2, hop 128).

    ncu --set full --clock-control none --import-source on -k regex:'stft256|istft' -s 4 -c 2 -o gpurun_out/x python scripts/prof_stft_only.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl4ss_b200 import features

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L, hop = 40000, 128
dev = torch.device('cuda:0')
wav = torch.randn(B, L, device=dev)
T = 1 + L // hop
masks = torch.rand(B, S, T, 129, device=dev)
for i in range(4):
    feat, cplx = features.stft_features(wav, 256, hop, 'hann', 'abs')
    out = features.mask_istft(masks, cplx, hop)
torch.cuda.synchronize()
print('ok', float(out.abs().mean()))

"""CPU oracle for the DL4SS separation hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in Python 3 / numpy / torch-CPU, what the reference
(`/root/reference`, Python 2 + librosa + mir_eval, cannot run in this image)
computes on the path  waveform -> STFT features -> BLSTM/BGRU encoder ->
speaker attention masks -> mask x mixture -> iSTFT (+ MSE loss).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker or as the timed
CPU baseline.  Nothing under `dl4ss_b200/` imports it: the product path has
no CPU fallback.

Parity status
-------------
* Model stages (a5-a11: nn.LSTM/nn.GRU/nn.Linear/nn.Embedding/baddbmm/
  sigmoid/tanh/MSELoss) run on the reference's own dependency (PyTorch, CPU
  fp32), so they ARE the reference arithmetic: pinned by construction.
* STFT/iSTFT (librosa, unpinned version, absent here) and SDR (mir_eval,
  absent here): **parity unpinned** against those libraries' values.  The
  restatements are pinned instead against independent implementations
  available here (`torch.stft/istft`, `scipy.signal`) and against the shape
  constants the reference embeds (39936, 313x129, 134x129) -- see
  tests/test_oracle_stft.py and tests/golden/.
"""

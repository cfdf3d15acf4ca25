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
* STFT/iSTFT: **pinned to the reference's own code.**  tests/golden/ref_stft_*.npz
  are outputs of the pure-numpy `sqrt_hann` / `stft` / `istft` the reference ships
  (Cocktail/software/DL4SS_Keras/test_stft_istft.py:9-63), produced by
  tests/golden/make_ref_fixtures.py, which exec()s those lines where they lie;
  oracle/stft_ref.py reproduces them to 1e-6 (tests/test_oracle_refpin.py) in
  the un-centred form those functions use and, on the shared frames / samples,
  in the centred librosa form the path implements.  librosa itself (version
  unpinned in the reference, absent here) is additionally cross-checked through
  `torch.stft/istft` and `scipy.signal`, and the shape constants the reference
  embeds (39936, 313x129, 134x129, 36480).
* SDR (mir_eval.separation, un-vendored, absent here): **parity unpinned**
  against that library's values; the restatement follows the published
  BSS-Eval v3 algorithm and is pinned by identities (tests/test_oracle.py).
"""

"""CPU tests of the host logic behind the reference-named entry points (dl4ss_b200/compat.py): list-file parsing,
dB rules, wav I/O, the synthetic source, and that nothing computes without the GPU."""
import os

import numpy as np
import pytest


def test_parse_mix_line_reference_format():
    from dl4ss_b200 import compat as c
    line = 'wsj0/si_tr_s/40n/40na010x.wav 1.9857 wsj0/si_tr_s/01x/01xc0203.wav -1.9857\n'
    assert c.parse_mix_line(line) == (['40n', '01x'], [1.9857, -1.9857], ['40na010x', '01xc0203'])
    three = 'a/011/011a0101.wav 0.5 b/20g/20ga010m.wav -1.25 c/40n/40na010x.wav 2 \n'
    spk, db, names = c.parse_mix_line(three)
    assert spk == ['011', '20g', '40n'] and db == [0.5, -1.25, 2.0] and names == ['011a0101', '20ga010m', '40na010x']


def test_three_speaker_db_rules():
    """Torch_multi/predata_multiAims_3dB.py:124-145: ranges of the drawn amplitude factors."""
    from dl4ss_b200 import compat as c
    rng = np.random.RandomState(0)
    hits = [0, 0]
    for _ in range(200):
        r = c.three_speaker_db_rates(2, 5, rng)
        assert sorted(r)[0] == 1.0 and 1.0 <= max(r) <= 10 ** 0.25
        hits[int(r[1] != 1.0)] += 1
        n, l, s = c.three_speaker_db_rates(3, 5, rng)
        assert abs(n - 10 ** 0.125) < 1e-12 and 10 ** 0.125 <= l <= 10 ** 0.25 and 1.0 <= s <= 10 ** 0.125
    assert min(hits) > 60                                   # either channel is picked about half of the time
    assert c.three_speaker_db_rates(2, 0, rng) == [1.0, 1.0]


def test_wav_io_and_list_source(tmp_path):
    from dl4ss_b200 import compat as c
    rng = np.random.RandomState(1)
    data, lists = tmp_path / 'data', tmp_path / 'lists'
    want = {}
    for spk, name in (('011', '011a0101'), ('20g', '20ga010m')):
        d = data / 'train' / spk
        d.mkdir(parents=True)
        x = rng.uniform(-0.9, 0.9, 3000)
        c.write_wav_pcm(str(d / (name + '.wav')), x, 8000)
        want[(spk, name)] = np.round(x * 32768.0) / 32768.0
    for split in ('eval', 'test'):
        (data / split).mkdir()
    lists.mkdir()
    (lists / 'mix_2_spk_cv.txt').write_text('x/011/011a0101.wav 1.5 y/20g/20ga010m.wav -1.5\n')
    src = c.ListFileSource(str(data), str(lists))
    assert sorted(src.speakers('train')) == ['011', '20g']
    rec = src.recipes('valid', 2)
    assert rec == [(['011', '20g'], [1.5, -1.5], ['011a0101', '20ga010m'])]
    for spk, name in zip(rec[0][0], rec[0][2]):
        x, rate = src.read('valid', spk, name)
        assert rate == 8000 and np.array_equal(x, want[(spk, name)])     # PCM16 round trip, exactly


def test_synthetic_source_is_deterministic():
    from dl4ss_b200 import compat as c
    a, b = c.SyntheticSource(4, seed=3), c.SyntheticSource(4, seed=3)
    ra = a.recipes('valid', 2)
    assert ra == b.recipes('valid', 2) and ra != a.recipes('train', 2)
    assert all(abs(d[0] + d[1]) < 1e-12 for _, d, _ in ra)              # WSJ0-2mix: +g / -g
    x, sr = a.read('valid', ra[0][0][0], ra[0][2][0])
    y, _ = b.read('valid', ra[0][0][0], ra[0][2][0])
    assert sr == 8000 and np.array_equal(x, y) and 2.5 * sr <= len(x) <= 6 * sr


def test_multi_label_vector_and_names():
    import dl4ss_b200 as d
    y_spk, y_map = d.multi_label_vector([{'b': 0, 'c': 0}, {'a': 0}], {'a': 0, 'b': 1, 'c': 2})
    assert y_spk == [[1, 2], [0]] and y_map.dtype == np.float32 and y_map.tolist() == [[0, 1, 1], [1, 0, 0]]
    for name in ('prepare_data', 'prepare_datasize', 'bss_eval', 'bss_eval_cRM', 'eval_bss'):
        assert callable(getattr(d, name))
    assert callable(d.bss_test.cal) and callable(d.predata_fromList.prepare_data)


def test_no_cpu_path():
    import torch
    import dl4ss_b200 as d
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        next(d.prepare_data('once', 'valid'))

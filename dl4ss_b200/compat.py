"""The reference's own entry points around the hot path, same names / arguments / return values, on the CUDA kernels.

    prepare_data(mode, train_or_test)      TDAA_beta/predata_fromList.py:45-234 ; cRM: TDAA_beta/predata_fromList_cRM_123.py:51-300
    prepare_datasize(gen)                  TDAA_beta/predata_fromList.py:36-43
    multi_label_vector(x, dict_name2idx)   TDAA_beta/test_multi_labels_speech.py:287-300
    top_k_mask(batch_pro, alpha, top_k)    TDAA_beta/main_run_sstune_EvalVer.py:390-405 (CPU FloatTensor result, as there)
    bss_eval(...) / bss_eval_cRM(...)      TDAA_beta/main_run_sstune_EvalVer.py:36-75 ; ...cRM_EvalVer.py:68-107
    eval_bss(...)                          TDAA_beta/main_run_sstune_EvalVer.py:407-507
    bss_test.cal(path, aim_mix_number)     Torch_multi/bss_test.py:12-61

What stays the reference's: the generator protocol ('global' tuple, 'once' dict keys, `False` at epoch end), the wav
file names under batch_output/, the PCM16 round trip before scoring, the evaluation loop's statements.  What runs on the
GPU instead of numpy/librosa/mir_eval: per-source preprocessing + mixing (`dl4ss_premix_shift_fwd`), every STFT of a
batch in two launches (`dl4ss_stft_feat`), the modules, mask x mixture + iSTFT (`dl4ss_mask_istft`), BSS-Eval
(`dl4ss_xcorr_f64` + Cholesky).  The dataset walk of the reference (os.listdir over a WSJ0 tree, soundfile, resampy)
is behind a small `Source` interface: `ListFileSource` parses the reference's WSJ0-2mix list files
(`path dB path dB` lines, predata_fromList.py:113-133), `SyntheticSource` makes WSJ0-2mix-shaped noise so that the
loops run without a dataset (none ships with the reference).
"""
import builtins
import os
import random
import re
import shutil
import wave
import zlib

import numpy as np
import torch

from . import config
from . import features
from . import metrics
from . import modules as M

# the reference's module-level switches (TDAA_beta/main_run_sstune_EvalVer.py:27-28)
test_all_outputchannel = 0
test_mode = 1


class _Lrs(object):
    """Stand-in for the reference's `lrs` logging service client: keeps what was sent, prints nothing."""

    def __init__(self):
        self.sent = []

    def send(self, *a):
        self.sent.append(a)


lrs = _Lrs()
log = print          # the reference prints progress; assign `compat.log = lambda *a: None` to silence


# ------------------------------------------------------------------------------------------------ data sources
def parse_mix_line(line):
    """One line of a create-speaker-mixtures list, `.../011/011a0101.wav 1.2 .../20g/20ga010m.wav -1.2`
    -> ([spk,...], [dB,...], [sample name,...]) with the reference's regular expressions
    (TDAA_beta/predata_fromList.py:113-116; they need the trailing separator the list files have)."""
    line = line if line.endswith(('\n', ' ')) else line + '\n'
    spk = re.findall('/([0-9][0-9].)/', line)
    db = [float(x) for x in re.findall(' (.*?) ', line.replace('\n', ' '))]
    names = re.findall(r'/(.{8})\.wav ', line)
    return spk, db, names


def read_wav_pcm(path):
    """Minimal PCM wav reader (stdlib `wave`; the reference uses soundfile) -> (float64 signal in [-1,1), rate)."""
    with wave.open(path, 'rb') as f:
        n, ch, width, rate = f.getnframes(), f.getnchannels(), f.getsampwidth(), f.getframerate()
        raw = f.readframes(n)
    if width != 2:
        raise ValueError('%s: only PCM16 wav files are read here' % path)
    x = np.frombuffer(raw, dtype='<i2').astype(np.float64) / 32768.0
    if ch > 1:
        x = x.reshape(-1, ch)[:, 0]                      # `signal = signal[:, 0]`, predata_fromList.py:132-133
    return x, rate


def write_wav_pcm(path, x, rate):
    """sf.write(path, x, rate) with soundfile's default WAV subtype (PCM16)."""
    q = np.round(np.clip(np.asarray(x, dtype=np.float64), -1.0, 1.0 - 2.0 ** -15) * 32768.0).astype('<i2')
    with wave.open(path, 'wb') as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(rate))
        f.writeframes(q.tobytes())


class Source(object):
    """What prepare_data needs from a dataset."""

    def speakers(self, split):
        """Speaker names of 'train' | 'eval' | 'test' (the directory listings of the reference)."""
        raise NotImplementedError

    def recipes(self, train_or_test, mix_k):
        """Ordered mixtures of the split: a list of ([spk...], [dB...], [sample name...])."""
        raise NotImplementedError

    def read(self, train_or_test, spk, sample_name):
        """-> (signal 1-D float, rate)."""
        raise NotImplementedError


class ListFileSource(Source):
    """The reference's WSJ0 layout: `data_path/{train,eval,test,eval_test}/<spk>/<sample>.wav` and
    `list_path/mix_{k}_spk_{tr,cv,tt}.txt` (TDAA_beta/predata_fromList.py:67-93,122-126)."""

    def __init__(self, data_path, list_path='./create-speaker-mixtures/', reader=read_wav_pcm):
        self.data_path, self.list_path, self.reader = data_path, list_path, reader

    def speakers(self, split):
        return os.listdir(os.path.join(self.data_path, split))

    def recipes(self, train_or_test, mix_k):
        suffix = {'train': 'tr', 'valid': 'cv', 'test': 'tt'}[train_or_test]
        with open(os.path.join(self.list_path, 'mix_{}_spk_{}.txt'.format(mix_k, suffix))) as f:
            return [parse_mix_line(l) for l in f.readlines()]

    def read(self, train_or_test, spk, sample_name):
        sub = 'train' if train_or_test != 'test' else 'eval_test'
        return self.reader(os.path.join(self.data_path, sub, spk, sample_name + '.wav'))


class SyntheticSource(Source):
    """WSJ0-2mix-shaped synthetic data: 101 training speakers, utterances of 2.5-6 s of band-shaped, syllable-gated
    noise at config.FRAME_RATE, gains uniform in +-2.5 dB with opposite signs (the WSJ0-2mix convention)."""

    def __init__(self, num_mixtures=64, num_spk=101, seed=1):
        self.n, self.seed = num_mixtures, seed
        self.spk = ['%02d%s' % (i // 26, 'abcdefghijklmnopqrstuvwxyz'[i % 26]) for i in range(num_spk)]

    def speakers(self, split):
        return list(self.spk) if split == 'train' else []

    def recipes(self, train_or_test, mix_k):
        rng = np.random.RandomState(self.seed + {'train': 0, 'valid': 1, 'test': 2}[train_or_test])
        out = []
        for i in range(self.n):
            spk = [self.spk[j] for j in rng.choice(len(self.spk), mix_k, replace=False)]
            g = rng.uniform(0.0, 2.5)
            db = [g, -g] + [float(rng.uniform(-2.5, 2.5)) for _ in range(mix_k - 2)]
            out.append((spk, db[:mix_k], ['%s%05d' % (s, i) for s in spk]))
        return out

    def read(self, train_or_test, spk, sample_name):
        rng = np.random.RandomState(zlib.crc32(('%s/%s/%d' % (spk, sample_name, self.seed)).encode()) % (2 ** 31))
        sr = config.FRAME_RATE
        n = int(rng.uniform(2.5, 6.0) * sr)
        t = np.arange(n) / float(sr)
        x = np.zeros(n)
        for _ in range(3):                               # three resonances of white noise
            f0, bw = rng.uniform(200.0, 3400.0), rng.uniform(80.0, 400.0)
            spec = np.fft.rfft(rng.standard_normal(n))
            fr = np.fft.rfftfreq(n, 1.0 / sr)
            x += np.fft.irfft(spec / (1.0 + ((fr - f0) / bw) ** 2), n) * rng.uniform(0.3, 1.0)
        env = 0.05 + 0.5 * (1.0 + np.sign(np.sin(2 * np.pi * rng.uniform(2.5, 5.0) * t + rng.uniform(0, 6.28))))
        return x * env + rng.uniform(-0.01, 0.01), sr


_source = None


def set_source(source):
    """Install the dataset `prepare_data` reads (a `Source`)."""
    global _source
    _source = source


def get_source():
    global _source
    if _source is None:
        _source = SyntheticSource()
    return _source


def three_speaker_db_rates(mix_k, dB=None, rng=np.random):
    """Per-channel amplitude factors of the generator that draws its own level differences
    (Torch_multi/predata_multiAims_3dB.py:124-145,190-210): two speakers -> one randomly chosen channel is scaled by
    10^(dB/20*u); three speakers -> channel 1 'normal' 10^(dB/40), channel 2 'large' 10^(dB/20*(0.5+0.5u)), channel 3
    'small' 10^(dB/20*0.5u).  Returns a list of mix_k factors (1.0 where the reference leaves a channel alone)."""
    dB = getattr(config, 'dB', 5) if dB is None else dB
    rates = [1.0] * mix_k
    if dB and mix_k == 2:
        rate = 10 ** (dB / 20.0 * rng.rand())
        rates[0 if rng.rand() > 0.5 else 1] = rate
    if dB and mix_k == 3:
        large = 10 ** (dB / 20.0 * (0.5 + 0.5 * rng.rand()))
        small = 10 ** (dB / 20.0 * (0.5 * rng.rand()))
        rates = [10 ** (dB / 20.0 * 0.5), large, small]
    return rates


# ------------------------------------------------------------------------------------------------ prepare_data
def prepare_datasize(gen):
    data = next(gen)
    return data[1].shape[1], data[1].shape[2], data[4].shape[1], data[-1], (data[4].shape[2], data[4].shape[3])


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('dl4ss_b200.compat: a CUDA device is required (no CPU path)')
    return torch.device('cuda', torch.cuda.current_device())


def prepare_data(mode, train_or_test, min=None, max=None, as_numpy=True):
    """Generator with the reference's protocol.

    mode 'global' yields (all_spk, dict_spk_to_idx, dict_idx_to_spk, T, F, 32, num_spk, batch_total) once;
    mode 'once' yields, per batch of config.BATCH_SIZE mixtures, the dict
        mix_wav [B,L] f64, mix_feas [B,T,F] f32, mix_phase [B,T,F] c64, (mix_mag [B,T,F,2] with config.is_ComlexMask),
        aim_fea [B,T,F(,2)], aim_spkname, query, num_all_spk, multi_spk_fea_list (B dicts spk -> [T,F] | [T,F,2]),
        multi_spk_wav_list (B dicts spk -> [L]), batch_total
    and `False` when the split's list is used up.  `min`/`max` override config.MIN_MIX / MAX_MIX (cRM_123 signature).
    as_numpy=False keeps the arrays as CUDA tensors (skips the device-to-host copies the reference loop then undoes
    with `torch.from_numpy(...).cuda()`)."""
    src = get_source()
    dev = _device()
    B, L = config.BATCH_SIZE, config.MAX_LEN
    n_fft, hop = config.FRAME_LENGTH, config.FRAME_SHIFT
    all_spk_train = src.speakers('train')
    all_spk = all_spk_train + src.speakers('eval') + src.speakers('test')
    mix_k = random.randint(config.MIN_MIX if min is None else min, config.MAX_MIX if max is None else max)
    all_samples_list = list(src.recipes(train_or_test, mix_k))
    batch_total = len(all_samples_list) // B
    if mode == 'global':
        spk = sorted(all_spk_train)
        yield (spk, {s: i for i, s in enumerate(spk)}, {i: s for i, s in enumerate(spk)},
               1 + L // hop, n_fft // 2 + 1, 32, len(spk), batch_total)
        return
    if mode != 'once':
        raise ValueError('Wrong input of mode.')
    if getattr(config, 'SHUFFLE_BATCH', False):
        random.shuffle(all_samples_list)
    augment = bool(getattr(config, 'AUGMENT_DATA', False)) and train_or_test == 'train'
    cplx = bool(config.is_ComlexMask)
    raw = torch.zeros(B, mix_k, L, dtype=torch.float32).pin_memory()
    for batch_idx in range(batch_total):
        raw.zero_()
        lengths = np.zeros((B, mix_k), np.int32)
        shifts = np.zeros((B, mix_k), np.int32)
        gains = np.zeros((B, mix_k), np.float32)
        names = []
        for b in range(B):
            spk_k, db_k, sample_k = all_samples_list[batch_idx * B + b]
            assert len(spk_k) == mix_k == len(db_k) == len(sample_k)
            names.append(spk_k)
            for k, spk in enumerate(spk_k):
                signal, rate = src.read(train_or_test, spk, sample_k[k])
                if rate != config.FRAME_RATE:
                    raise ValueError('sample rate %d != config.FRAME_RATE: resampling is outside this path' % rate)
                n = builtins.min(len(signal), L)
                raw[b, k, :n] = torch.from_numpy(np.asarray(signal[:n], dtype=np.float32))
                lengths[b, k], gains[b, k] = n, db_k[k]
                if augment:
                    shifts[b, k] = random.sample(range(n), 1)[0]
        pm = features.premix(raw.to(dev, non_blocking=True), torch.from_numpy(gains), torch.from_numpy(lengths),
                             shifts=torch.from_numpy(shifts) if augment else None)
        batch = features.prepare_batch(pm['mix_wav'], n_fft, hop, sources=pm['sources'])
        tgt = batch['multi_spk_mag'] if cplx else batch['multi_spk_fea']           # [B,S,T,F(,2)]
        conv = (lambda t: t.cpu().numpy()) if as_numpy else (lambda t: t)
        mix_wav = conv(pm['mix_wav'].double())
        tgt_h, src_h = conv(tgt), conv(pm['sources'].double() if as_numpy else pm['sources'])
        out = {'mix_wav': mix_wav,
               'mix_feas': conv(batch['mix_feas']),
               'mix_phase': conv(batch['mix_phase']),
               'aim_fea': tgt_h[:, 0],
               'aim_spkname': [n[0] for n in names],
               'query': np.array([]),
               'num_all_spk': len(all_spk),
               'multi_spk_fea_list': [{spk: tgt_h[b, k] for k, spk in enumerate(names[b])} for b in range(B)],
               'multi_spk_wav_list': [{spk: src_h[b, k] for k, spk in enumerate(names[b])} for b in range(B)],
               'batch_total': batch_total}
        if cplx:
            out['mix_mag'] = conv(batch['mix_mag'])
        yield out
    yield False


def multi_label_vector(x, dict_name2idx):
    y_spk, y_aim = [], []
    length = len(dict_name2idx)
    for sample in x:
        tmp_vector = [0 for _ in range(length)]
        line = []
        for spk in sample.keys():
            line.append(dict_name2idx[spk])
            for l in line:
                tmp_vector[l] = 1
        y_spk.append(line)
        y_aim.append(tmp_vector)
    return y_spk, np.array(y_aim, dtype=np.float32)


def top_k_mask(batch_pro, alpha, top_k):
    """The reference returns a CPU FloatTensor (its caller does `.numpy()` on it)."""
    return M.top_k_mask(batch_pro, alpha, top_k).cpu()


def print_spk_name(d, batch):
    return ' '.join(str([d[int(j)] for j in i]) for i in batch)


# ------------------------------------------------------------------------------------------------ bss_eval
def _as_cuda(x, dtype=None):
    t = x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))
    t = t.detach().to(_device())
    return t if dtype is None else t.to(dtype)


class SepBatch(object):
    """What one bss_eval call produced, kept in memory: per mixture the speaker names (file-name order) and the
    waveforms `pre`, `genTrue`, `realTrue` [n_spk, n] + `mix` [n] as CUDA float32 tensors."""

    def __init__(self):
        self.samples = []

    def add(self, spk, pre, gen, real, mix):
        self.samples.append({'spk': spk, 'pre': pre, 'genTrue': gen, 'realTrue': real, 'mix': mix})


last_batch = None        # the SepBatch of the most recent bss_eval / bss_eval_cRM call (bss_test.cal(None, n) scores it)


def _bss_eval_common(pred_spec, true_spec, y_map_gtruth, dict_idx2spk, train_data, dst='batch_output'):
    """pred_spec / true_spec: complex spectra [B,S,T,F,2] on the device -> waveforms (one K6 launch each), wav files."""
    global last_batch
    write = bool(config.Out_Sep_Result)
    if write:
        if os.path.exists(dst):
            log(" \ncleanup: " + dst + "/")
            shutil.rmtree(dst)
        os.makedirs(dst)
    hop = config.FRAME_SHIFT
    wav_pre = features.mask_istft(None, pred_spec.contiguous(), hop)               # librosa.istft(spec.T, hop) per (b, s)
    wav_gen = features.mask_istft(None, true_spec.contiguous(), hop)
    n_pre = wav_pre.shape[-1]
    out = SepBatch()
    rate = config.FRAME_RATE
    for sample_idx, each_sample in enumerate(train_data['multi_spk_wav_list']):
        chans = [(idx, dict_idx2spk[int(one_cha)]) for idx, one_cha in enumerate(y_map_gtruth[sample_idx]) if one_cha]
        min_len = n_pre                                 # test_mode / test_all_outputchannel: len(wav_pre)
        if not test_mode and not test_all_outputchannel:
            min_len = builtins.min([n_pre] + [len(each_sample[s]) for _, s in chans])
        pre = torch.stack([wav_pre[sample_idx, idx, :min_len] for idx, _ in chans])
        gen = torch.stack([wav_gen[sample_idx, idx, :min_len] for idx, _ in chans])
        real_names = sorted(each_sample.keys())
        real = torch.stack([_as_cuda(each_sample[s], torch.float32)[:39936] for s in real_names])
        mix = _as_cuda(train_data['mix_wav'][sample_idx], torch.float32)[:min_len]
        out.add({'pre': [s for _, s in chans], 'realTrue': real_names}, pre, gen, real, mix)
        if write:
            for k, s in enumerate(real_names):
                write_wav_pcm('{}/{}_{}_realTrue.wav'.format(dst, sample_idx, s), real[k].cpu().numpy(), rate)
            for k, (_, s) in enumerate(chans):
                write_wav_pcm('{}/{}_{}_pre.wav'.format(dst, sample_idx, s), pre[k].cpu().numpy(), rate)
                write_wav_pcm('{}/{}_{}_genTrue.wav'.format(dst, sample_idx, s), gen[k].cpu().numpy(), rate)
            write_wav_pcm('{}/{}_True_mix.wav'.format(dst, sample_idx), mix.cpu().numpy(), rate)
    last_batch = out
    return out


def _unit_phase(mix_phase):
    """exp(1j*np.angle(X)) as [B,T,F,2] (angle(0) = 0 -> 1+0j, as numpy)."""
    x = _as_cuda(mix_phase)
    x = torch.view_as_real(x) if torch.is_complex(x) else x
    x = x.to(torch.float32)
    mag = torch.sqrt(x[..., 0] ** 2 + x[..., 1] ** 2)
    unit = x / mag.clamp_min(1e-30).unsqueeze(-1)
    unit[..., 0] = torch.where(mag > 0, unit[..., 0], torch.ones_like(mag))
    return unit


def bss_eval(predict_multi_map, y_multi_map, y_map_gtruth, dict_idx2spk, train_data):
    """Reconstruct `pred * exp(1j*angle(mix))` and the same for the targets, write batch_output/*.wav (when
    config.Out_Sep_Result), return the waveforms (`SepBatch`).  y_map_gtruth: per mixture the speaker indices of the
    output channels (the reference passes top_k_mask_idx; a zero entry reads as 'channel off', as there)."""
    unit = _unit_phase(train_data['mix_phase']).unsqueeze(1)                       # [B,1,T,F,2]
    pred = _as_cuda(predict_multi_map, torch.float32).unsqueeze(-1) * unit
    true = _as_cuda(y_multi_map, torch.float32).unsqueeze(-1) * unit
    return _bss_eval_common(pred, true, y_map_gtruth, dict_idx2spk, train_data)


def bss_eval_cRM(predict_map_real, predict_map_fake, y_multi_map, y_map_gtruth, dict_idx2spk, train_data):
    pred = torch.stack([_as_cuda(predict_map_real, torch.float32), _as_cuda(predict_map_fake, torch.float32)], -1)
    return _bss_eval_common(pred, _as_cuda(y_multi_map, torch.float32), y_map_gtruth, dict_idx2spk, train_data)


class _BssTest(object):
    """`import bss_test; bss_test.cal(path, aim_mix_number)` (Torch_multi/bss_test.py:12-61)."""

    add_slience_channel = 0

    @staticmethod
    def _score(aim, pre):
        if pre.shape[0] == 1 and aim.shape[0] == 2:
            pre = pre.repeat(2, 1)
        n = builtins.min(aim.shape[-1], pre.shape[-1])
        sdr = metrics.bss_eval_sources_batch(aim[None, :, :n].contiguous(), pre[None, :, :n].contiguous())[0]
        return sdr[0].cpu().numpy()

    def cal(self, path, aim_mix_number=2, pcm16=True):
        """SDR of every mixture under `path` (the wav files bss_eval wrote), or of an in-memory `SepBatch` (None: the
        last bss_eval call's; pcm16=True applies the quantisation the files would have)."""
        SDR_sum = np.array([])
        if path is None or isinstance(path, SepBatch):
            batch = last_batch if path is None else path
            q = metrics.pcm16_roundtrip if pcm16 else (lambda x: x)
            for s in batch.samples:
                order = np.argsort(s['spk']['pre'])                       # sorted(os.listdir): speaker-name order
                SDR_sum = np.append(SDR_sum, self._score(q(s['realTrue']), q(s['pre'][list(order)])))
            log('SDR here:', SDR_sum.mean())
            return SDR_sum
        files = [l for l in sorted(os.listdir(path)) if l[-3:] == 'wav']
        mix_number = len(set(l.split('_')[0] for l in files))
        log('num of mixed :', mix_number)
        dev = _device()
        for idx in range(mix_number):
            pre, aim = [], []
            for l in files:
                if l.split('_')[0] == str(idx):
                    if 'realTrue' in l:
                        aim.append(read_wav_pcm(os.path.join(path, l))[0])
                    if 'pre' in l:
                        pre.append(read_wav_pcm(os.path.join(path, l))[0])
            aim = torch.from_numpy(np.array(aim, dtype=np.float32)).to(dev)
            pre = torch.from_numpy(np.array(pre, dtype=np.float32)).to(dev)
            SDR_sum = np.append(SDR_sum, self._score(aim, pre))
        log('SDR here:', SDR_sum.mean())
        return SDR_sum


bss_test = _BssTest()


# ------------------------------------------------------------------------------------------------ eval_bss
def Variable(x, requires_grad=False):
    return x


def eval_bss(mix_hidden_layer_3d, adjust_layer, mix_speech_classifier, mix_speech_multiEmbedding, att_speech_layer,
             loss_multi_func, dict_spk2idx, dict_idx2spk, num_labels, mix_speech_len, speech_fre):
    """The reference's evaluation epoch, statement for statement (python 3 syntax; `lrs`, `bss_test`, `prepare_data`,
    `top_k_mask`, `bss_eval` are this module's).  Returns SDR_SUM (the reference only prints / sends its mean)."""
    for i in [mix_speech_multiEmbedding, adjust_layer, mix_speech_classifier, mix_hidden_layer_3d, att_speech_layer]:
        i.training = False
    log('#' * 40)
    eval_data_gen = prepare_data('once', 'valid')
    SDR_SUM = np.array([])
    with torch.no_grad():
        while True:
            eval_data = next(eval_data_gen)
            if eval_data is False:
                break
            mix_speech_hidden, mix_tmp_hidden = mix_hidden_layer_3d(Variable(torch.from_numpy(eval_data['mix_feas'])).cuda())
            mix_speech_output = mix_speech_classifier(Variable(torch.from_numpy(eval_data['mix_feas'])).cuda())

            if not test_mode:
                y_spk_list = eval_data['multi_spk_fea_list']
                y_spk_gtruth, y_map_gtruth = multi_label_vector(y_spk_list, dict_spk2idx)
                if not test_mode and getattr(config, 'Ground_truth', True):
                    mix_speech_output = Variable(torch.from_numpy(y_map_gtruth)).cuda()
                    if test_all_outputchannel:
                        mix_speech_output = Variable(torch.ones(config.BATCH_SIZE, num_labels, ))
                        y_map_gtruth = np.ones([config.BATCH_SIZE, num_labels])

            if test_mode:
                num_labels = 2
                alpha0 = -0.5
            else:
                alpha0 = 0.5
            top_k_mask_mixspeech = top_k_mask(mix_speech_output, alpha=alpha0, top_k=num_labels)
            top_k_mask_idx = [np.where(line == 1)[0] for line in top_k_mask_mixspeech.numpy()]
            log('Predict spk list:', print_spk_name(dict_idx2spk, top_k_mask_idx))
            mix_speech_multiEmbs = mix_speech_multiEmbedding(top_k_mask_mixspeech, top_k_mask_idx)
            mix_adjust = adjust_layer(mix_tmp_hidden, mix_speech_multiEmbs)
            mix_speech_multiEmbs = mix_adjust + mix_speech_multiEmbs

            assert len(top_k_mask_idx[0]) == len(top_k_mask_idx[-1])
            top_k_num = len(top_k_mask_idx[0])

            mix_speech_hidden_5d = mix_speech_hidden.view(config.BATCH_SIZE, 1, mix_speech_len, speech_fre, config.EMBEDDING_SIZE)
            mix_speech_hidden_5d = mix_speech_hidden_5d.expand(config.BATCH_SIZE, top_k_num, mix_speech_len, speech_fre, config.EMBEDDING_SIZE).contiguous()
            mix_speech_hidden_5d_last = mix_speech_hidden_5d.view(-1, mix_speech_len, speech_fre, config.EMBEDDING_SIZE)
            att_multi_speech = att_speech_layer(mix_speech_hidden_5d_last, mix_speech_multiEmbs.view(-1, config.EMBEDDING_SIZE))
            att_multi_speech = att_multi_speech.view(config.BATCH_SIZE, top_k_num, mix_speech_len, speech_fre)
            multi_mask = att_multi_speech

            x_input_map = Variable(torch.from_numpy(eval_data['mix_feas'])).cuda()
            x_input_map_multi = x_input_map.view(config.BATCH_SIZE, 1, mix_speech_len, speech_fre).expand(config.BATCH_SIZE, top_k_num, mix_speech_len, speech_fre)
            predict_multi_map = multi_mask * x_input_map_multi

            y_multi_map = np.zeros([config.BATCH_SIZE, top_k_num, mix_speech_len, speech_fre], dtype=np.float32)
            batch_spk_multi_dict = eval_data['multi_spk_fea_list']
            if test_mode:
                for iiii in range(config.BATCH_SIZE):
                    y_multi_map[iiii] = np.array(list(batch_spk_multi_dict[iiii].values()))
            else:
                for idx, sample in enumerate(batch_spk_multi_dict):
                    y_idx = sorted([dict_spk2idx[spk] for spk in sample.keys()])
                    if not test_mode:
                        assert y_idx == list(top_k_mask_idx[idx])
                    for jdx, oo in enumerate(y_idx):
                        y_multi_map[idx, jdx] = sample[dict_idx2spk[oo]]
            y_multi_map = Variable(torch.from_numpy(y_multi_map)).cuda()

            loss_multi_speech = loss_multi_func(predict_multi_map, y_multi_map)

            y_sum_map = Variable(torch.ones(config.BATCH_SIZE, mix_speech_len, speech_fre)).cuda()
            predict_sum_map = torch.sum(multi_mask, 1)
            loss_multi_sum_speech = loss_multi_func(predict_sum_map, y_sum_map)
            log('loss 1 eval, losssum eval : ', loss_multi_speech.data.cpu().numpy(), loss_multi_sum_speech.data.cpu().numpy())
            lrs.send('loss mask eval:', loss_multi_speech.data.cpu().item())
            lrs.send('loss sum eval:', loss_multi_sum_speech.data.cpu().item())
            loss_multi_speech = loss_multi_speech + 0.5 * loss_multi_sum_speech
            log('evaling multi-abs norm this eval batch:', torch.abs(y_multi_map - predict_multi_map).norm().data.cpu().numpy())
            log('loss:', loss_multi_speech.data.cpu().numpy())
            bss_eval(predict_multi_map, y_multi_map, top_k_mask_idx, dict_idx2spk, eval_data)
            SDR_SUM = np.append(SDR_SUM, bss_test.cal('batch_output/' if config.Out_Sep_Result else None, 2))
            log('SDR_aver_now:', SDR_SUM.mean())

    SDR_aver = SDR_SUM.mean() if SDR_SUM.size else float('nan')
    log('SDR_SUM (len:{}) for epoch eval : '.format(SDR_SUM.shape))
    lrs.send('SDR eval aver', SDR_aver)
    log('#' * 40)
    return SDR_SUM

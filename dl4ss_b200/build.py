"""Build libdl4ss_b200.so in-tree with nvcc for sm_100a (no torch, no pybind: plain C ABI).

    python -m dl4ss_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libdl4ss_b200.so')
SOURCES = ['api.cu', 'stft.cu', 'gemm.cu', 'gemm_tc.cu', 'attn.cu', 'rnn.cu', 'rnn_tc.cu', 'rnn_mma.cu', 'train.cu', 'rnn_bwd.cu', 'xcorr.cu', 'disc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Wno-deprecated-gpu-targets',
              '--use_fast_math=false', '-Xcompiler', '-fPIC']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, '..', 'include', 'dl4ss_b200.h'), os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.o')
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false'] + ['-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(out.decode())
        if pr.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [_nvcc(), '-shared', '-o', OUT] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

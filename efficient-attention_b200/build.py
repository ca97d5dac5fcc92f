"""Build libeva_sm100.so in-tree (efficient-attention_b200/lib/) with nvcc for sm_100a.

    python efficient-attention_b200/build.py [--force]

No torch / pybind dependency: the library is plain CUDA behind the C ABI of include/eva_sm100.h.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIB_DIR, 'libeva_sm100.so')
SOURCES = ['abi.cu', 'eva_generic.cu', 'eva_backward.cu', 'eva_window_tc_sm100.cu', 'eva_bwd_sm100.cu', 'eva_window_bwd_gen_sm100.cu', 'lara_generic.cu', 'lara_backward.cu', 'rfa_kernels.cu', 'rfa_tc_sm100.cu', 'sb_window_tc_sm100.cu', 'ra_sample_tc_sm100.cu', 'eva_fused_sm100.cu', 'eva_cluster_sm100.cu', 'eva_causal_sm100.cu', 'lara_core_sm100.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '--use_fast_math=false', '-Xptxas', '-v']


def _nvcc():
    cand = os.environ.get('NVCC') or os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'bin', 'nvcc')
    return cand if os.path.exists(cand) else 'nvcc'


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.h')):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, '.build_digest')
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    flags = [f for f in NVCC_FLAGS if f != '--use_fast_math=false']
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [_nvcc()] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'==== {src} ====\n{out}')
        failed |= p.returncode != 0
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if failed:
        sys.stderr.write('\n'.join(log))
        raise RuntimeError('nvcc failed; see efficient-attention_b200/lib/build.log')
    link = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    open(stamp, 'w').write(digest)
    if verbose:
        print(f'built {LIB}')
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)

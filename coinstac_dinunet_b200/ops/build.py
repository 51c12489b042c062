"""In-tree build of the sm_100a kernel library ``_b200_ops.so`` with plain nvcc.

``python -m coinstac_dinunet_b200.ops.build`` (or ``__graft_entry__.build()``).  Every ``csrc/*.cu`` is
compiled with ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (nvcc cross-compiles without a
GPU) and linked into one shared object that lives next to this file, so it travels with the repo
snapshot to the GPU box.  The library exposes a C ABI (``coinn_*``) consumed through ctypes by
``ops/native.py`` - no torch headers are involved, a full rebuild takes well under a minute.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, '_b200_ops.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
BASE_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
              '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']
FLAGS = BASE_FLAGS + ['--use_fast_math']
# Optimizer arithmetic is IEEE (sqrt / divide of the Adam update match torch.optim bit for bit up to summation order);
# fast-math stays on for the conv / GEMM / normalisation kernels whose epilogues tolerate approximate rsqrt / exp.
EXACT_MATH = {'fused_reduce_opt.cu', 'powersgd.cu', 'lowrank.cu'}


def flags_for(src):
    return BASE_FLAGS if os.path.basename(src) in EXACT_MATH else FLAGS


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path):
    h = hashlib.sha1()
    for p in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))):
        with open(p, 'rb') as fp:
            h.update(fp.read())
    h.update(' '.join(ARCH + flags_for(path)).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
    stamp = obj + '.sha1'
    dig = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC, *ARCH, *flags_for(src), '-I', CSRC, '-c', src, '-o', obj]
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    with open(stamp, 'w') as fp:
        fp.write(dig)
    return obj, True


LAST_BUILD = {}      # what the most recent build() in this process did (also written to profiles/build_record.json)


def build(verbose=False, force=False):
    """Compile every csrc/*.cu for sm_100a and link ``_b200_ops.so``.  Objects are cached by the sha1 of (source, headers,
    flags); ``force=True`` or ``COINN_FORCE_REBUILD=1`` recompiles everything from scratch.  What happened - per source
    "compiled" or "cached", the nvcc version, the flags, the sha256 of the library - is recorded in ``LAST_BUILD``."""
    import time
    os.makedirs(OBJ_DIR, exist_ok=True)
    force = force or os.environ.get('COINN_FORCE_REBUILD') == '1'
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = sources()
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs) or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    relinked = any(changed for _, changed in results) or not os.path.exists(LIB)
    if relinked:
        cmd = [NVCC, *ARCH, '-shared', '-o', LIB, *objs, '-lcudart']
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    try:
        ver = subprocess.run([NVCC, '--version'], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    except Exception:
        ver = 'unknown'
    with open(LIB, 'rb') as fp:
        lib_sha = hashlib.sha256(fp.read()).hexdigest()
    LAST_BUILD.clear()
    LAST_BUILD.update({
        'mode': 'from_scratch' if all(c for _, c in results) else ('incremental' if relinked else 'cached'),
        'forced': bool(force), 'seconds': round(time.time() - t0, 2), 'nvcc': ver, 'arch': ARCH[1],
        'flags': {'default': FLAGS, 'exact_math_files': sorted(EXACT_MATH)},
        'sources': {os.path.basename(s): ('compiled' if c else 'cached') for s, (_, c) in zip(srcs, results)},
        'relinked': bool(relinked), 'library': os.path.relpath(LIB, os.path.dirname(os.path.dirname(HERE))),
        'library_sha256': lib_sha, 'library_bytes': os.path.getsize(LIB)})
    return LIB


if __name__ == '__main__':
    print(build(verbose='-q' not in sys.argv, force='--force' in sys.argv))

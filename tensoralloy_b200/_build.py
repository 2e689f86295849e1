"""Compile libtab200.so in-tree with nvcc for sm_100a (no JIT cache: the built
.so travels to the GPU box with the repo snapshot)."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / 'tensoralloy_b200' / 'csrc'
SOURCES = ['scan.cu', 'nbr.cu', 'eam.cu', 'sf.cu', 'hessian.cu', 'pairs.cu', 'td_heads.cu', 'dd.cu']
# -cudart shared: the library carries no private copy of the CUDA runtime (and none of its
# entry-point tables); it uses the libcudart.so.12 already loaded by torch / the system one
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-cudart', 'shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return 'nvcc'


def needs_build():
    out = CSRC / 'libtab200.so'
    if not out.exists():
        return True
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob('*.h')) + \
        list(CSRC.glob('*.cuh')) + [ROOT / 'include' / 'tab200.h']
    return any(d.stat().st_mtime > out.stat().st_mtime for d in deps)


def build_library(force=False, verbose=False, defines=(), out_name='libtab200.so'):
    """`defines` / `out_name`: A/B variants of the kernels (tools/var_bench.sh), selected at
    run time with TAB200_LIB."""
    out = CSRC / out_name
    if not force and out_name == 'libtab200.so' and not needs_build():
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + [f'-D{d}' for d in defines] + \
        [f'-I{ROOT / "include"}', f'-I{CSRC}', '-o', str(out)] + [str(CSRC / s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out

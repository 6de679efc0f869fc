"""Build libffgp.so for sm_100a with nvcc (no torch headers: the boundary is a plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'libffgp.so')
SOURCES = ['dense_gp.cu', 'kron.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-shared',
         '-Xcompiler', '-fPIC', '-Xptxas=-v']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(HERE, '..', '..', 'include', 'ffgp.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = ['nvcc'] + FLAGS + ['-o', OUT] + [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed building libffgp.so')
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print(OUT)

"""Build the C part of the oracle (oracle/_build/liboracle_chaos.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'liboracle_chaos.so')


def build(force=False):
    src = os.path.join(HERE, 'chaos.c')
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-shared', '-fPIC',
           '-o', LIB, src, '-lm']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('oracle build failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force=True))

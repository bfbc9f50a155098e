"""In-tree build of libpb2.so (sm_100a only) with plain nvcc.

Usage: python -m probability_b200.build [--force]
The shared object lands in probability_b200/_C/libpb2.so; it is git-ignored but
travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB = os.path.join(OUT_DIR, 'libpb2.so')
SOURCES = ['pb2_chain_kernels.cu', 'pb2_misc.cu', 'pb2_capi.cu', 'pb2_rowshard.cu', 'pb2_dense_tc.cu', 'pb2_tile.cu', 'pb2_tile_nuts.cu', 'pb2_logistic_tc.cu', 'pb2_comm.cu', 'pb2_user.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']


def _nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return 'nvcc'


def _deps():
  deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
  deps.append(os.path.join(HERE, '..', 'include', 'pb2.h'))
  return deps


def needs_build():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
  """Compile every CUDA source for sm_100a and link libpb2.so. Returns the path."""
  os.makedirs(OUT_DIR, exist_ok=True)
  if not force and not needs_build():
    return LIB
  nvcc = _nvcc()
  objs = []
  procs = []
  for src in SOURCES:
    path = os.path.join(CSRC, src)
    if not os.path.exists(path):
      continue
    obj = os.path.join(OUT_DIR, src.replace('.cu', '.o'))
    cmd = ([nvcc] + NVCC_FLAGS + os.environ.get('PB2_NVCC_EXTRA', '').split() +
           (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj])
    procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs.append(obj)
  for cmd, pr in procs:
    out, _ = pr.communicate()
    if verbose or pr.returncode != 0:
      sys.stderr.write(out.decode())
    if pr.returncode != 0:
      raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
  link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-ldl']
  subprocess.check_call(link)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

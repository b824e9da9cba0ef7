"""N>1 path on the CPU: world_size-2 gloo run of the sharded batched driver + diagnostics gather must give the
same per-chain traces as a single process (chains are independent units; results do not depend on the number
of ranks)."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def run(world, out_path, port):
    worker = os.path.join(ROOT, 'tests', 'dist_worker.py')
    env = dict(os.environ, OPENBLAS_NUM_THREADS='1')
    if world == 1:
        cmd = [sys.executable, worker, out_path]
    else:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
               '--master-addr', '127.0.0.1', '--master-port', str(port), worker, out_path]
    subprocess.run(cmd, check=True, env=env, timeout=600, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return np.load(out_path)


def test_two_ranks_equal_one(tmp_path):
    a = run(1, str(tmp_path / 'w1.npz'), 0)
    b = run(2, str(tmp_path / 'w2.npz'), 29533)
    assert a['thetas'].shape == (5, 25, 2)
    assert np.array_equal(a['thetas'], b['thetas'])
    assert np.array_equal(a['n_cubic_ops'], b['n_cubic_ops'])
    assert np.all(np.isfinite(b['thetas']))

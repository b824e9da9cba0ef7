"""Worker for tests/test_distributed_cpu.py: run with torch.distributed (gloo) at world_size W; every rank
runs its shard of chains with the CPU oracle backend and the diagnostics are all-gathered."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)


def main(out_path):
    import torch.distributed as dist
    from apm_b200 import batched, synth
    from apm_b200.distributed import shard_chains, gather_diagnostics
    from oracle_backend import OracleBackend
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1:
        dist.init_process_group('gloo')
    n_chains, n_iter, N = 5, 25, 2
    X, y, _ = synth.make_dataset(40, 2, seed=11)
    mine = shard_chains(n_chains, rank, world)
    seeds = [1000 + c for c in mine]
    out = {'thetas': np.zeros((0, n_iter, 2)), 'n_cubic_ops': np.zeros((0,))}
    if mine:
        drv = batched.BatchedAPMSampler(OracleBackend(X, y), 40, N, 2, 'ess+rdss', batched.make_log_prior(2, False), seeds)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = drv.get_samples(None, n_iter, theta_init_sampler=lambda prng: synth.draw_theta_prior(prng, 2, ard=False))
        out = {'thetas': res['thetas'], 'n_cubic_ops': res['n_cubic_ops'].astype(float)}
    full = gather_diagnostics(out, n_chains, mine)
    if rank == 0:
        np.savez(out_path, **full)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])

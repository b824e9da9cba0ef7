"""Multi-GPU plumbing: independent chains are sharded over ranks (one process per GPU) with NO collective on
the estimator's critical path; the only exchange is a gather of per-chain diagnostics.

Chain c always uses seed `seed_base + c` and the same data set, so a chain's trace does not depend on the
number of ranks.  torch.distributed is used for the plumbing: backend "nccl" on GPUs (NVLink / NVSwitch),
"gloo" in the CPU tests."""
import numpy as np


def shard_chains(n_chains, rank, world_size):
    """Contiguous block partition of chain indices [0, n_chains) for `rank`."""
    base, rem = divmod(n_chains, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def gather_diagnostics(local, n_chains, chain_ids, device=None):
    """All-gather per-chain arrays.  `local` maps name -> array whose first axis runs over this rank's chains
    (`chain_ids`); returns name -> array over all n_chains chains (same on every rank).  Messages are
    KB-to-MB sized and off the critical path (one call at the end of a run / every k iterations)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = {}
        for k, v in local.items():
            v = np.asarray(v)
            full = np.full((n_chains,) + v.shape[1:], np.nan, dtype=np.float64)
            full[np.asarray(chain_ids, dtype=int)] = v
            out[k] = full
        return out
    world = dist.get_world_size()
    counts = [len(shard_chains(n_chains, r, world)) for r in range(world)]
    max_count = max(counts)
    out = {}
    for k in sorted(local):
        v = np.asarray(local[k], dtype=np.float64)
        tail = v.shape[1:]
        buf = torch.full((max_count,) + tail, float('nan'), dtype=torch.float64)
        if v.shape[0]:
            buf[:v.shape[0]] = torch.from_numpy(np.ascontiguousarray(v))
        if device is not None:
            buf = buf.to(device)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        full = np.concatenate([p.cpu().numpy()[:counts[r]] for r, p in enumerate(parts)], axis=0)
        out[k] = full
    return out

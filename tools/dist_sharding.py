"""Multi-GPU use of the decode path: shots are independent, so each rank decodes a contiguous slice on its own GPU
(with its own replica of the decoder tables) and the only exchange is one sum of counters and one gather of the
predictions. This mirrors how the reference scales out -- sinter gives every worker process its own compiled decoder
and adds up the per-process error counts (/root/reference/README.md:83-90, tools/step2_collect_stats.sh:21-29) -- with
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) as the plumbing.
"""
import numpy as np


def shard_bounds(num_shots, rank, world_size):
    """Contiguous, balanced shot range [lo, hi) of `rank`; the first (num_shots % world_size) ranks get one extra."""
    base, extra = divmod(int(num_shots), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def decode_sharded(predict, dets, actual_obs=None, dist=None, device=None):
    """Decodes `dets` (the FULL batch, identical on every rank) data-parallel over the ranks of `dist`.

    predict: callable(dets_slice uint8[n, nbytes]) -> uint8[n, obs_bytes]   (this rank's decoder)
    Returns (predictions for the full batch uint8[num_shots, obs_bytes], total mistakes or None), identical on every rank.
    With dist=None (or a world of 1) this is just predict(dets).
    """
    import torch
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n = dets.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    local = np.ascontiguousarray(predict(dets[lo:hi]))
    ob = local.shape[1]
    mistakes = None
    if actual_obs is not None:
        got = np.zeros(hi - lo, np.uint64)
        for k in range(ob):
            got |= local[:, k].astype(np.uint64) << np.uint64(8 * k)
        mistakes = int((got != np.asarray(actual_obs[lo:hi], np.uint64)).sum())
    if world == 1:
        return local, mistakes
    dev = device if device is not None else torch.device('cpu')
    # predictions: ranks may hold slices of different length -> pad to the longest, gather, trim
    longest = shard_bounds(n, 0, world)[1]
    padded = torch.zeros((longest, ob), dtype=torch.uint8, device=dev)
    padded[:hi - lo] = torch.from_numpy(local).to(dev)
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n, r, world)
        parts.append(gathered[r][:b - a].cpu().numpy())
    full = np.concatenate(parts, axis=0)
    if mistakes is not None:
        t = torch.tensor([mistakes], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        mistakes = int(t.item())
    return full, mistakes

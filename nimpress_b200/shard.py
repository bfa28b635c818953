"""Multi-GPU plumbing for the scoring path (one process per GPU, torch.distributed).

Two ways the path shards:

* variant-sharded (default): rank r scores a contiguous range of the score rows over ALL samples.
  Counts and decisions are local; the only exchange is the final combine of the per-sample partial
  sums: all-gather, then add in rank order on every rank (fixed order => every rank gets the same
  bits, independent of the collective's algorithm), plus an integer sum of nloci.
* sample-sharded: rank r holds a slab of samples for every row.  The per-row tallies must be summed
  across ranks BEFORE the decision: one integer all-reduce of [n_rows, 2] per block (exact, order
  free); afterwards nothing floating-point crosses ranks -- scores are concatenated.

Works with the nccl backend on device tensors and with gloo on CPU tensors (tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, world, rank):
    """Contiguous balanced range [lo, hi) of rank `rank`; sizes differ by at most one."""
    q, r = divmod(n_items, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def combine_partials(sums, nloci, group=None):
    """Variant-sharded combine.  sums: float64 [n_samples] partial sums of this rank; nloci: int64
    tensor [1].  Returns (total_sums, total_nloci) identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return sums.clone(), nloci.clone()
    parts = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(parts, sums.contiguous(), group=group)
    total = parts[0].clone()
    for p in parts[1:]:                       # rank order = score-file order of the ranges
        total += p
    nl = nloci.clone()
    dist.all_reduce(nl, op=dist.ReduceOp.SUM, group=group)
    return total, nl


def combine_counts(counts, group=None):
    """Sample-sharded: sum the per-row integer tallies [n_rows, 2] (nmiss, neff) over ranks, in place."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_scores(scores, sizes, group=None):
    """Sample-sharded: concatenate the ranks' score slabs (sizes[r] samples on rank r)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return scores.clone()
    m = max(sizes)
    pad = torch.zeros(m, dtype=scores.dtype, device=scores.device)
    pad[:scores.numel()] = scores
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])

"""Sharding of reference views over GPUs.  Reference views are independent units (the
reference runs one process per view, scripts/pipes.sh:30-49), so the only parallel strategy is data-parallel
over views with NO collective on the data path; an optional gather of the per-view results feeds fusion."""


def views_for_rank(n_views, rank, world):
    """Round-robin assignment r -> GPU r mod world (balanced to within one view)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_views, world))


def block_for_rank(n_views, rank, world):
    """Contiguous assignment: rank r gets views [start, start + count) with counts differing by at most one.  Views that
    are neighbours in the list (capture order) share most of their pair.txt source views, so a rank decodes and keeps
    resident little more than its own block -- what the multi-view driver uses (cli.run_all_views)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_results(local, world, group=None):
    """Optional terminal gather of {view_id: ndarray} dicts to rank 0 (torch.distributed object gather; gloo or
    nccl).  Not on the hot path."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    bucket = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if bucket is None:
        return None
    out = {}
    for part in bucket:
        out.update(part)
    return out


def run_sharded(n_views, rank, world, process_view):
    """Calls process_view(view_id) for this rank's views; returns {view_id: result}."""
    return {v: process_view(v) for v in views_for_rank(n_views, rank, world)}

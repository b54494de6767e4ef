"""Multi-GPU layout of the batch: images are independent, so the index range is
cut into contiguous shards, one per rank (one process per GPU), each rank runs the whole
pipeline on its shard, and the only exchange is a host-side gather of the polygon lists in
image order.  No data-path collective exists or is needed."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [first, first+count) of rank's shard; the first n_items % world ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_polygon_scores(polygons, scores, dst=0, group=None):
    """polygons / scores: this rank's per-image lists (PolygonScores fields).  Returns the
    concatenation over ranks in rank (= image index) order on `dst`, None elsewhere.
    Works on any torch.distributed backend (the payload is a few KB per image)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(polygons), list(scores)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = (list(polygons), list(scores))
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst, group=group)
    if rank != dst:
        return None
    all_p, all_s = [], []
    for p, s in out:
        all_p.extend(p)
        all_s.extend(s)
    return all_p, all_s


def gather_polygons(result, dst=0, group=None):
    """result: this rank's _ffi.Polygons (flat arrays).  Returns the _ffi.Polygons of the whole
    batch on `dst` (shards concatenated in rank = image index order), None elsewhere.  The
    payload is the five flat arrays, a few bytes per polygon point."""
    import torch.distributed as dist
    from ._ffi import Polygons
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return result
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = [None] * world if rank == dst else None
    dist.gather_object(result.arrays(), out, dst=dst, group=group)
    return Polygons.concat(out) if rank == dst else None

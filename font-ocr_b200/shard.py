"""Page sharding across GPUs.

Pages are independent units (the reference already maps one rayon task per page, ncc.rs:839-846,
main.rs:443-467, and re-sorts by page index, ncc.rs:847, main.rs:468), so the multi-GPU path is: one
process per GPU, a contiguous block of pages per rank, the template / glyph bank replicated to every
GPU once, NO data-path collective, and a host-side gather of the per-page results ordered by page
index.  `torch.distributed` is only plumbing here (rendezvous, barrier, the final object gather).
"""
from __future__ import annotations


def shard_range(n_pages: int, rank: int, world: int) -> range:
    """Contiguous block of page indices for `rank`: sizes differ by at most one, lower ranks first."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_pages, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_by_page(local: dict, group=None, dst: int = 0):
    """Gather {page_index: result} from every rank onto `dst` and return the results as a list ordered
    by page index (ncc.rs:847 `pages.sort_by_key(|(i, _)| *i)`); other ranks get None.  Works on any
    backend (gloo objects on CPU, NCCL groups fall back to a gloo side group made by the caller)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local[i] for i in sorted(local)]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged = {}
    for part in bucket:
        dup = set(part) & set(merged)
        if dup:
            raise ValueError(f"pages {sorted(dup)} were processed by more than one rank")
        merged.update(part)
    return [merged[i] for i in sorted(merged)]

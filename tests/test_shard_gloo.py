"""The N>1 path on CPU: world_size-2 gloo, page sharding without a data-path collective and the
host-side gather ordered by page index."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pages, out_q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from font_ocr_b200 import ncc, shard

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = shard.shard_range(n_pages, rank, world)
    # stand-in for the scan: each page's "result" is post-processed hits -> text, through the real
    # host-side post-processing (ncc.process_hits) so the gathered objects are what a run produces
    local = {}
    for p in mine:
        hits = [(chr(65 + (p + k) % 26), 10 + 16 * k, 7, 0.96) for k in range(5)]
        local[p] = ncc.lines_to_text(ncc.process_hits(hits))
    res = shard.gather_by_page(local)
    dist.barrier()
    if rank == 0:
        out_q.put(res)
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from font_ocr_b200 import shard

    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                rr = shard.shard_range(n, r, w)
                seen += list(rr)
                assert len(rr) in (n // w, n // w + 1)
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_two_rank_gather_by_page_index():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    n_pages = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
    [p.start() for p in procs]
    res = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert len(res) == n_pages
    for p, lines in enumerate(res):
        assert lines == ["".join(chr(65 + (p + k) % 26) for k in range(5))]

"""CPU suite, part 3: the N>1 path (page sharding + scatter/gather plumbing) with world_size 2
over gloo.  The per-page work is a CPU stand-in (the oracle's Sauvola): what is tested is that
sharded results, gathered, are byte-identical to the single-process result in page order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_pages, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from archive_pdf_tools_b200 import shard, synth
    from oracle import oracle as orc
    H, W = 60, 80
    pages = None
    if rank == 0:
        pages = torch.from_numpy(np.stack([synth.make_page(i, H, W, dpi=100, rgb=False) for i in range(n_pages)]))
    local = shard.scatter_pages(pages, (H, W), n_pages, src=0)
    assert local.shape[0] == len(shard.shard_indices(n_pages, rank, world))
    masks = torch.from_numpy(np.stack([orc.sauvola(p.numpy(), 25).view(np.uint8) for p in local]) if local.shape[0]
                             else np.zeros((0, H, W), np.uint8))
    full = shard.gather_results(masks, n_pages, dst=0)
    t = shard.max_over_ranks(1.0 + rank)
    if rank == 0:
        exp = np.stack([orc.sauvola(p.numpy(), 25).view(np.uint8) for p in pages])
        q.put((bool(np.array_equal(full.numpy(), exp)), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_pages', [5, 1])
def test_sharded_pipeline_matches_single_process(n_pages):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and t == 2.0


def test_shard_indices_partition():
    from archive_pdf_tools_b200 import shard
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 4, 8):
            parts = [shard.shard_indices(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert shard.shard_counts(n, w) == [len(p) for p in parts]

#!/usr/bin/env python
"""One book through the multi-GPU data path, one process per GPU under torchrun:

    root synthesises P pages -> shard.scatter_pages (NCCL send/recv, page i -> rank i mod G) -> every rank runs
    b200mrc_decompose on its shard -> shard.gather_results (NCCL) -> root hashes mask / fg / bg in page order.

Prints ONE JSON line on rank 0: SHA-256 of the gathered planes (must not depend on the world size) and the two
throughput figures the survey asks for -- "pre-sharded" (compute only, max over ranks) and "root-scatter-included"
(scatter + compute + gather, wall clock of the slowest rank).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/shard_run.py --pages 32 --config 3
"""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=32)
    ap.add_argument('--config', type=int, default=3)
    ap.add_argument('--shape', type=int, nargs=2, default=None, help='override H W of the config (small test pages)')
    ap.add_argument('--reps', type=int, default=3, help='timed repetitions of scatter -> compute -> gather')
    a = ap.parse_args()
    import bench
    cfg = dict(bench.CONFIGS[a.config])
    if a.shape:
        cfg['H'], cfg['W'] = a.shape
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    P, H, W, C = a.pages, cfg['H'], cfg['W'], cfg['C']
    pages_host = None
    if rank == 0:                                        # before CUDA is touched: the generator forks worker processes
        ncpu = len(os.sched_getaffinity(0))
        distinct = bench.make_pages(cfg, 0, min(P, 16), max(1, min(16, ncpu // 2)))
        pages_host = np.stack([distinct[(i * 7) % len(distinct)] for i in range(P)])

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29517')
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local_rank))
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import shard
    dev = torch.device('cuda', local_rank)
    eng = pkg.get_engine()
    shape = (H, W, C) if C == 3 else (H, W)
    pages_dev = torch.from_numpy(pages_host).to(dev) if rank == 0 else None
    n_local = len(shard.shard_indices(P, rank, world))
    batch = eng.make_batch(max(n_local, 1), H, W, C, bg_downsample=cfg['bg'], mask_only=cfg['mask_only'])

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        sync()
        t0 = time.time()
        local = shard.scatter_pages(pages_dev, shape, P, src=0, device=dev)
        torch.cuda.synchronize()
        t1 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if n_local:
            batch.img.upload(local)                      # device -> pitched device plane
            e0.record()
            batch.run(cfg['window'], denoise_mask='fast')
            e1.record()
        torch.cuda.synchronize()
        t2 = time.time()
        outs = {}
        for name in (('mask',) if cfg['mask_only'] else ('mask', 'fg', 'bg')):
            plane = getattr(batch, name)
            loc = plane.view()[:n_local].contiguous()
            outs[name] = shard.gather_results(loc, P, dst=0)
        torch.cuda.synchronize()
        t3 = time.time()
        comp_ms = e0.elapsed_time(e1) if n_local else 0.0
        return outs, (t1 - t0, comp_ms / 1e3, t3 - t2, t3 - t0)

    outs, _ = one_pass()                                 # warm-up (NCCL communicators, workspaces)
    times = []
    for _ in range(a.reps):
        outs, t = one_pass()
        times.append([shard.max_over_ranks(v, device=dev) for v in t])
    if rank == 0:
        best = min(times, key=lambda t: t[3])
        px = P * H * W
        hashes = {k: hashlib.sha256(v.cpu().numpy().tobytes()).hexdigest() for k, v in outs.items()}
        print(json.dumps({
            'world': world, 'pages': P, 'page': [H, W, C], 'config': cfg['name'], 'window': cfg['window'],
            'sha256': hashes,
            'scatter_s': best[0], 'compute_s': best[1], 'gather_s': best[2], 'total_s': best[3],
            'pre_sharded_Mpx_s': px / best[1] / 1e6 if best[1] > 0 else None,
            'root_scatter_included_Mpx_s': px / best[3] / 1e6,
            'scatter_GB_s': P * H * W * C * (world - 1) / world / best[0] / 1e9 if world > 1 else None,
            'note': 'times are the max over ranks; pre-sharded = b200mrc_decompose only (CUDA events), root-scatter-included = '
                    'NCCL scatter from the root GPU + decompose + NCCL gather of mask/fg/bg to the root GPU (wall clock)'}),
            flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Host<->device copy ceilings of this box, one process per GPU (plain python for one GPU, torchrun for N): pinned 1-D
copies each way and both ways at once, all ranks copying concurrently after a barrier.  Rank 0 prints one JSON line with
the per-GPU and the aggregate rates (max time over ranks) -- the ceiling the end-to-end path (bench.py `e2e`) runs against.

  python tools/pcie_bw.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/pcie_bw.py
"""
import json, os, sys, time
import torch
import torch.distributed as dist


def main():
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=5):
        fn(); sync(); t0 = time.time()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.time() - t0) / reps], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        sync()
        return float(dt.item())

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    a = timed(lambda: d.copy_(h, non_blocking=True))
    b = timed(lambda: h2.copy_(d2, non_blocking=True))
    c = timed(both)
    if rank == 0:
        gb = n / 1e9
        print(json.dumps({'n_gpus': world, 'bytes_per_copy': n,
                          'h2d_GB_s_per_gpu': gb / a, 'd2h_GB_s_per_gpu': gb / b, 'duplex_GB_s_per_gpu_each_way': gb / c,
                          'h2d_GB_s_aggregate': world * gb / a, 'd2h_GB_s_aggregate': world * gb / b,
                          'duplex_GB_s_aggregate_total': 2 * world * gb / c,
                          'cpus': len(os.sched_getaffinity(0)),
                          'note': 'all ranks copy concurrently; times are the max over ranks'}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

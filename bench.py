#!/usr/bin/env python
"""bench.py -- MRC decompose throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a batch of 64 RGB pages 3300x2550 @400 DPI per GPU, full
create_mrc_hocr_components (window 101, denoise 'fast', bg_downsample=3), synthetic pages.
A step = one pass of the whole path over the batch.  `value` = Mpixels/s with the batch resident
in HBM (CUDA events, max over ranks); `e2e` = the same through the public host API
(archive_pdf_tools_b200.decompose_pages: pinned host buffers -> H2D -> b200mrc_decompose -> D2H).
Pages shard over ranks with no data-path collective ("weak" scaling: 64 pages per GPU).
--impl reference times the reference's own CPU implementation (oracle/_ref Cython + Pillow +
scipy) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W, C, DPI, WINDOW, BG_DS, PAGES_PER_GPU = 3300, 2550, 3, 400, 101, 3, 64
DISTINCT_PAGES = 16                    # generated per rank; tiled to the 64-page batch
METRIC = 'MRC decompose Mpixels/sec @400-DPI pages (64 RGB pages 3300x2550 per GPU, bg/3, denoise fast)'
WORKLOAD = 'configs[1]: batch 64 RGB pages 3300x2550 @400 DPI, full MRC decompose, bg-downsample=3'
BYTES_PER_PX_PIPELINE = 3 + 1 + 3 + 3.0 / 9           # 7.333 B/px
BYTES_PER_PX_OPTIMISE = 3 + 1 + 3 + 3                 # optimise stage (k_opt_fir_w + k_opt_iir_w): img + mask in, fg + bg out


def _gen_page(idx):
    import archive_pdf_tools_b200.synth as synth
    return synth.make_page(idx, H, W, dpi=DPI, rgb=True, sigma_n=3.0, halftone=False)


def make_pages(first, count, workers):
    from concurrent.futures import ProcessPoolExecutor
    idxs = list(range(first, first + count))
    if workers <= 1:
        return [_gen_page(i) for i in idxs]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_gen_page, idxs))


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.12):
                continue
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- reference (CPU) arm
def _ref_init(kind):
    """Worker start-up (untimed): load the reference kernels and synthesise this worker's page."""
    global _decomp, _page
    import archive_pdf_tools_b200.synth as synth
    if kind == 'reference':
        from oracle import ref_pipeline as rp
        rp.ref_modules()
        _decomp = rp.ref_decompose
    else:
        from oracle import oracle as orc
        orc.lib()
        _decomp = orc.decompose
    _page = synth.make_page(os.getpid() % 4096, H, W, dpi=DPI, rgb=True, sigma_n=3.0)


def _ref_page(_):
    t = time.time()
    res = _decomp(_page, dpi=DPI, bg_downsample=BG_DS, denoise_mask='fast')
    return time.time() - t, int(res['mask'].sum())


def cpu_pool():
    """(pool, cores, kind): all host cores, reference Cython when oracle/_ref exists else the C port."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import build_ref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    kind = 'reference' if build_ref.have_ref() else 'port'
    pool = ProcessPoolExecutor(max_workers=cores, initializer=_ref_init, initargs=(kind,))
    return pool, cores, kind


def cpu_step(pool, cores, kind, pages_per_core=1, first=0):
    """One bounded CPU sample: pages_per_core pages on every core; returns (Mpx/s, n_pages, seconds)."""
    n = cores * pages_per_core
    t0 = time.time()
    list(pool.map(_ref_page, range(first, first + n)))
    dt = time.time() - t0
    return n * H * W / dt / 1e6, n, dt


def run_reference(args, rank):
    if rank != 0:
        return 0
    pool, cores, kind = cpu_pool()
    _, _, t1 = cpu_step(pool, cores, kind)            # pool start-up + imports + page synthesis, untimed
    _, _, t1 = cpu_step(pool, cores, kind)            # a real step with every core busy
    # bounded sample: keep K timed steps within ~2.5 minutes by using fewer workers if necessary
    used = cores
    budget = 150.0
    if t1 * args.steps > budget:
        used = max(min(cores, 8), int(cores * budget / (t1 * args.steps)))
        pool.shutdown()
        from concurrent.futures import ProcessPoolExecutor
        pool = ProcessPoolExecutor(max_workers=used, initializer=_ref_init, initargs=(kind,))
        cpu_step(pool, used, kind)
    for _ in range(max(0, min(args.warmup, 2) - 1)):
        cpu_step(pool, used, kind)
    t0 = time.time()
    npages = 0
    for s in range(args.steps):
        _, n, _ = cpu_step(pool, used, kind)
        npages += n
    dt = time.time() - t0
    pool.shutdown()
    val = npages * H * W / dt / 1e6
    sample = '%d pages per step (1 per worker process, %d of %d host cores busy) of the 64-page batch' % (used, used, cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'Mpixels/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'pages_per_step': used, 'page': [H, W, C], 'window': WINDOW,
                   'bg_downsample': BG_DS, 'denoise': 'fast'},
        'cpu_baseline': {'value': val, 'unit': 'Mpixels/s', 'cores': used, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'Mpixels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-pipelined', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 1)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return run_reference(args, rank)

    # ---- synthetic pages first (forks worker processes: before CUDA is touched)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    gen_workers = max(1, min(DISTINCT_PAGES, ncpu // max(world, 1)))
    distinct = make_pages(rank * DISTINCT_PAGES, DISTINCT_PAGES, gen_workers)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib

    N = PAGES_PER_GPU
    host = torch.empty((N, H, W, C), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for i in range(N):
        hv[i] = distinct[i % DISTINCT_PAGES]
    del distinct
    eng = pkg.get_engine()
    batch = eng.make_batch(N, H, W, C, bg_downsample=BG_DS)
    batch.img.upload(host, non_blocking=False)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps (inputs in HBM, 1.6 GB per batch >> 126 MB L2: no flush needed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                              # nvidia-smi needs a moment to start: begin before the warm-up
    for _ in range(args.warmup):
        batch.run_staged(WINDOW, denoise_mask='fast')
    torch.cuda.synchronize()
    t_w = time.time()
    while rank == 0 and sampler.proc and not sampler.rows and time.time() - t_w < 3.0:
        batch.run_staged(WINDOW, denoise_mask='fast')   # extra untimed warm-up until the sampler is alive
        torch.cuda.synchronize()
    barrier()
    launches0 = _lib.lib().b200mrc_launch_count()
    _lib.profile_enable(True)                        # CUDA-event pair around every kernel of the timed region
    stage_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts0 = sampler.mark()
    e0.record()
    for _ in range(args.steps):
        ev = {}
        batch.run_staged(WINDOW, denoise_mask='fast', events=ev)
        stage_events.append(ev)
    e1.record()
    barrier()
    ts1 = sampler.mark()
    launches = _lib.lib().b200mrc_launch_count() - launches0
    kernel_ms = {k: v[1] / max(v[0], 1) for k, v in _lib.profile_report().items()}   # mean ms per launch
    _lib.profile_enable(False)
    clocks = sampler.stop(ts0, ts1) if rank == 0 else None
    dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in [ev[k] for ev in stage_events if k in ev]]))
                for k in stage_events[0]}
    px_step = N * H * W * world
    value = px_step * args.steps / (dev_ms / 1e3) / 1e6

    # ---- the same K steps with two batches in flight (one CUDA stream per in-flight batch):
    # the row-latency-bound sweep of one batch overlaps the throughput-bound stages of the next.  Reported next
    # to `value`, which stays the plain one-batch-at-a-time figure the per-kernel numbers belong to.
    pipelined = None
    if not args.no_pipelined:
        batch2 = eng.make_batch(N, H, W, C, bg_downsample=BG_DS)
        batch2.img = batch.img                          # same resident input pages, private outputs / workspace
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        both = [batch, batch2]

        def pipelined_steps(k):
            cur = torch.cuda.current_stream()
            for s_ in streams:
                s_.wait_stream(cur)
            for i in range(k):
                with torch.cuda.stream(streams[i % 2]):
                    both[i % 2].run(WINDOW, denoise_mask='fast')      # b200mrc_decompose: private workspace per batch
            for s_ in streams:
                cur.wait_stream(s_)

        pipelined_steps(4)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pipelined_steps(args.steps)
        p1.record()
        barrier()
        pt = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        pipelined = {'value': px_step * args.steps / (float(pt.item()) / 1e3) / 1e6, 'unit': 'Mpixels/s',
                     'ms_per_step': float(pt.item()) / args.steps, 'batches_in_flight': 2}
        del batch2, both

    # ---- end to end through the public host API: pinned host pages -> results in pinned host memory
    e2e = None
    if not args.no_e2e:
        from archive_pdf_tools_b200.engine import StreamedDecomposer
        del batch
        torch.cuda.empty_cache()
        e2e_chunk = int(os.environ.get('B200MRC_E2E_CHUNK', '4'))
        e2e_streams = int(os.environ.get('B200MRC_E2E_STREAMS', '2'))
        e2e_buffers = int(os.environ.get('B200MRC_E2E_BUFFERS', '4'))
        sd = StreamedDecomposer(eng, N, H, W, C, chunk=e2e_chunk, bg_downsample=BG_DS, buffers=e2e_buffers,
                                compute_streams=e2e_streams)
        # a book is a stream of batches: two batches in flight (each with its own pinned result buffers), so the H2D
        # copies and kernels of step i+1 overlap the D2H tail of step i; every step's inputs cross PCIe and every
        # step's results are waited for and read on the host inside the timed region
        outs2 = [sd.alloc_outputs() for _ in range(2)]
        outs = outs2[0]

        def e2e_run(steps, depth):
            pending = []
            seen = 0
            for i in range(steps):
                o = outs2[i % 2]
                pending.append((sd.run_async(host, o, WINDOW, denoise_mask='fast'), o))
                if len(pending) >= depth:
                    ev, oo = pending.pop(0)
                    ev.synchronize()                              # mask/fg/bg of that step are in pinned host memory
                    seen += int(oo['mask'][0, 0, 0]) + int(oo['bg'][-1, -1, -1])
            for ev, oo in pending:
                ev.synchronize()
                seen += int(oo['mask'][0, 0, 0]) + int(oo['bg'][-1, -1, -1])
            return seen

        e2e_run(2, 2)
        e2e_steps = max(3, min(args.steps, 10))
        torch.cuda.synchronize()
        barrier()
        t0 = time.time()
        e2e_run(e2e_steps, 1)                                     # one call at a time (run() semantics)
        torch.cuda.synchronize()
        barrier()
        dt_sync = torch.tensor([time.time() - t0], dtype=torch.float64, device='cuda')
        t0 = time.time()
        e2e_run(e2e_steps, 2)
        torch.cuda.synchronize()
        barrier()
        dt = torch.tensor([time.time() - t0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dt_sync, op=dist.ReduceOp.MAX)
        e2e = {'value': px_step * e2e_steps / float(dt.item()) / 1e6, 'unit': 'Mpixels/s', 'steps': e2e_steps,
               'h2d_bytes_per_step': int(host.numel()) * world,
               'd2h_bytes_per_step': int(sum(v.numel() for v in outs.values())) * world,
               'one_call_at_a_time': px_step * e2e_steps / float(dt_sync.item()) / 1e6, 'batches_in_flight': 2,
               'api': 'archive_pdf_tools_b200.engine.StreamedDecomposer.run_async: pinned host pages -> 1-D H2D DMA -> device pitching '
                      '(b200mrc_copy2d) -> b200mrc_decompose (%d-page chunks, %d compute streams, %d device buffers) -> device '
                      'unpitching -> 1-D D2H DMA of mask/fg/bg into pinned host buffers' % (e2e_chunk, e2e_streams, e2e_buffers)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s'
    # dominant kernel of the step = the largest mean launch time in the timed region
    dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else None
    dom_ms = kernel_ms.get(dom)
    achieved = BYTES_PER_PX_OPTIMISE * N * H * W / (dom_ms / 1e3) / 1e9 if dom_ms else None
    opt_ms = stage_ms.get('optimise')
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak if achieved else None, 'traffic': None, 'peak_source': peak_src,
                'algorithmic_bytes_per_px': BYTES_PER_PX_OPTIMISE, 'kernel_ms': dom_ms,
                'note': 'algorithmic bytes = the optimise stage (img 3 + mask 1 in, fg 3 + bg 3 out) x pixels per launch; '
                        'the stage is k_opt_fir_w + k_opt_iir_w (the sweep that completes it is the dominant kernel), frac_stage uses both',
                'frac_stage': (BYTES_PER_PX_OPTIMISE * N * H * W / (opt_ms / 1e3) / 1e9 / peak) if opt_ms else None,
                'pipeline_frac': BYTES_PER_PX_PIPELINE * N * H * W / (dev_ms / args.steps / 1e3) / 1e9 / peak,
                'kernel_ms_all': kernel_ms, 'stage_ms': stage_ms}
    try:
        roofline['traffic'] = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(dom)
    except Exception:
        pass

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        pool, cores, kind = cpu_pool()
        cpu_step(pool, cores, kind)                       # start-up, untimed
        v, n, dt = cpu_step(pool, cores, kind, pages_per_core=2, first=1000)
        pool.shutdown()
        cpu_baseline = {'value': v, 'unit': 'Mpixels/s', 'cores': cores, 'kind': kind,
                        'sample': '%d pages (2 per core) of the workload in %.1f s wall; reference Cython (oracle/_ref) + '
                                  'Pillow + scipy' % (n, dt)}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'Mpixels/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8', 'data': 'synthetic (%d distinct pages per GPU tiled to %d)' % (DISTINCT_PAGES, N),
        'config': {'workload': WORKLOAD, 'pages_per_gpu': N, 'page': [H, W, C], 'window': WINDOW, 'k': 0.34,
                   'bg_downsample': BG_DS, 'denoise': 'fast', 'parallelism': 'pages sharded over %d GPU(s), no collective' % world,
                   'l2': 'inputs (1.6 GB/batch) larger than L2'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'pipelined': pipelined,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

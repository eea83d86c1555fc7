#!/usr/bin/env python
"""bench.py -- MRC decompose throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[1]: a batch of 64 RGB pages 3300x2550 @400 DPI per GPU, full
create_mrc_hocr_components (window 101, denoise 'fast', bg_downsample=3), synthetic pages.  --config picks one of the
five BASELINE configs (table CONFIGS below); the default line also carries short runs of configs
3, 4 and 5 under `extra_configs`.
A step = one pass of the whole path over the batch = ONE b200mrc_decompose call.  `value` = Mpixels/s with the batch
resident in HBM (CUDA events, max over ranks); `e2e` = the same through the public host API (pinned host buffers -> H2D
-> b200mrc_decompose -> D2H).  Pages shard over ranks with no data-path collective ("weak" scaling).
--impl reference times the reference's own CPU implementation (oracle/_ref Cython + Pillow + scipy) on all host cores,
on a bounded sample of the same workload.
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

# BASELINE.json configs -> concrete synthetic workloads.  bytes_px = algorithmic HBM bytes per
# input pixel: every input read once, every reference-visible output written once.
CONFIGS = {
    1: dict(name='configs[0]: single 3300x2550 grayscale page, Sauvola window=33 (binarise_sauvola-level call)',
            H=3300, W=2550, C=1, dpi=132, window=33, bg=None, mask_only=True, sauvola_only=True, halftone=0.0,
            pages=1, distinct=1, bytes_px=2.0),
    2: dict(name='configs[1]: batch 64 RGB pages 3300x2550 @400 DPI, full MRC decompose, bg-downsample=3',
            H=3300, W=2550, C=3, dpi=400, window=101, bg=3, mask_only=False, sauvola_only=False, halftone=0.0,
            pages=64, distinct=16, bytes_px=3 + 1 + 3 + 3.0 / 9),
    3: dict(name='configs[2]: 1000-page book @300 DPI (3300x2550 RGB, 30 % halftone pages), 125 pages per GPU and step',
            H=3300, W=2550, C=3, dpi=300, window=75, bg=3, mask_only=False, sauvola_only=False, halftone=0.3,
            pages=125, distinct=20, bytes_px=3 + 1 + 3 + 3.0 / 9),
    4: dict(name='configs[3]: 600-DPI scans 5100x6600 RGB, Sauvola window=51, denoise on, bg-downsample=3',
            H=6600, W=5100, C=3, dpi=600, window=51, bg=3, mask_only=False, sauvola_only=False, halftone=0.0,
            pages=16, distinct=4, bytes_px=3 + 1 + 3 + 3.0 / 9),
    5: dict(name='configs[4]: mask-only (1-bit) mode, gray pages 2200x1700 @200 DPI, window 51, denoise fast',
            H=2200, W=1700, C=1, dpi=200, window=51, bg=None, mask_only=True, sauvola_only=False, halftone=0.0,
            pages=256, distinct=16, bytes_px=2.0),
}
METRIC = 'MRC decompose Mpixels/sec @400-DPI pages (64 RGB pages 3300x2550 per GPU, bg/3, denoise fast)'

# Algorithmic bytes per input pixel of each kernel of the full-decompose step (C = channels):
# what the kernel must read and write if nothing were re-read -- the denominators of roofline.per_kernel.
def kernel_alg_bytes_px(C, bg):
    return {
        'k_noise_dd_march': 0.25 * C,                 # the centre crop (1/4 of the page) once
        'k_noise_select': 0.25,                       # one 32-bit key per 2x2 crop pixels... read once
        'k_sauvola_fused': C + 1.0,                   # page in, mask out (gray conversion and pre-blur fused)
        'k_gray_blur_fast': C + 1.0, 'k_sauvola_mask': 2.0,
        'k_mask_denoise': 1.0,                        # mask in (sparse write-back)
        'k_opt_fir_w': C + 1.0 + 8.0,                 # page + mask in, record plane out
        'k_opt_iir_w': 8.0 + C + 2.0 * C,             # records + page in, fg + bg out
        'k_resample_tile': C + C / float(bg * bg) if bg else 0.0,
    }


def _gen_page(args):
    idx, H, W, C, dpi, halftone = args
    import archive_pdf_tools_b200.synth as synth
    return synth.make_page(idx, H, W, dpi=dpi, rgb=(C == 3), sigma_n=3.0, halftone=halftone)


def make_pages(cfg, first, count, workers):
    from concurrent.futures import ProcessPoolExecutor
    n_ht = int(round(cfg['halftone'] * count))
    jobs = [(first + i, cfg['H'], cfg['W'], cfg['C'], cfg['dpi'], i < n_ht) for i in range(count)]
    if workers <= 1:
        return [_gen_page(j) for j in jobs]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_gen_page, jobs))


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
def _ref_init(kind, cfg):
    """Worker start-up (untimed): load the reference kernels and synthesise this worker's page."""
    global _decomp, _page, _cfg
    import archive_pdf_tools_b200.synth as synth
    _cfg = cfg
    if kind == 'reference':
        from oracle import ref_pipeline as rp
        rp.ref_modules()
        _decomp = rp
    else:
        from oracle import oracle as orc
        orc.lib()
        _decomp = orc
    idx = os.getpid() % 4096
    _page = synth.make_page(idx, cfg['H'], cfg['W'], dpi=cfg['dpi'], rgb=(cfg['C'] == 3), sigma_n=3.0,
                            halftone=(idx % 10) < int(round(cfg['halftone'] * 10)))


def _ref_page(_):
    t = time.time()
    cfg = _cfg
    if cfg['sauvola_only']:
        if hasattr(_decomp, 'ref_threshold_image'):
            m = _decomp.ref_threshold_image(_page, None, window=cfg['window'])
        else:
            m = _decomp.sauvola(_page, cfg['window'])
        return time.time() - t, int(m.sum())
    fn = _decomp.ref_decompose if hasattr(_decomp, 'ref_decompose') else _decomp.decompose
    res = fn(_page, dpi=cfg['dpi'], window=cfg['window'], bg_downsample=cfg['bg'], denoise_mask='fast', mask_only=cfg['mask_only'])
    return time.time() - t, int(res['mask'].sum())


def cpu_pool(cfg, workers=None):
    """(pool, cores, kind): all host cores, reference Cython when oracle/_ref exists else the C port."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import build_ref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    if workers:
        cores = min(cores, workers)
    kind = 'reference' if build_ref.have_ref() else 'port'
    pool = ProcessPoolExecutor(max_workers=cores, initializer=_ref_init, initargs=(kind, cfg))
    return pool, cores, kind


def cpu_step(pool, cores, cfg, pages_per_core=1):
    """One bounded CPU sample: pages_per_core pages on every core; returns (Mpx/s, n_pages, seconds)."""
    n = cores * pages_per_core
    t0 = time.time()
    list(pool.map(_ref_page, range(n)))
    dt = time.time() - t0
    return n * cfg['H'] * cfg['W'] / dt / 1e6, n, dt


def cpu_baseline_sample(cfg):
    """~10-30 s of CPU work on all host cores: 2 pages per core of the workload (reported next to the GPU numbers)."""
    pool, cores, kind = cpu_pool(cfg)
    cpu_step(pool, cores, cfg)                       # start-up, untimed
    ppc = 2 if cfg['H'] * cfg['W'] < 2e7 else 1
    v, n, dt = cpu_step(pool, cores, cfg, pages_per_core=ppc)
    pool.shutdown()
    return {'value': v, 'unit': 'Mpixels/s', 'cores': cores, 'kind': kind,
            'sample': '%d pages (%d per core) of the workload in %.1f s wall; reference Cython (oracle/_ref) + Pillow + scipy'
                      % (n, ppc, dt)}


def run_reference(args, rank, cfg):
    if rank != 0:
        return 0
    pool, cores, kind = cpu_pool(cfg)
    cpu_step(pool, cores, cfg)                       # pool start-up + imports + page synthesis, untimed
    _, _, t1 = cpu_step(pool, cores, cfg)            # a real step with every core busy
    # bounded sample: keep K timed steps within ~2.5 minutes by using fewer workers if necessary
    used = cores
    budget = 150.0
    if t1 * args.steps > budget:
        used = max(min(cores, 8), int(cores * budget / (t1 * args.steps)))
        pool.shutdown()
        pool, used, kind = cpu_pool(cfg, workers=used)
        cpu_step(pool, used, cfg)
    for _ in range(max(0, min(args.warmup, 2) - 1)):
        cpu_step(pool, used, cfg)
    t0 = time.time()
    npages = 0
    for s in range(args.steps):
        _, n, _ = cpu_step(pool, used, cfg)
        npages += n
    dt = time.time() - t0
    pool.shutdown()
    val = npages * cfg['H'] * cfg['W'] / dt / 1e6
    sample = '%d pages per step (1 per worker process, %d of %d host cores busy) of the %d-page batch' % (used, used, cores, cfg['pages'])
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'Mpixels/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': bench_config(cfg, args.gpus, sampled_pages_per_step=used),
        'cpu_baseline': {'value': val, 'unit': 'Mpixels/s', 'cores': used, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'Mpixels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def bench_config(cfg, world, **extra):
    c = {'workload': cfg['name'], 'pages_per_gpu': cfg['pages'], 'page': [cfg['H'], cfg['W'], cfg['C']], 'window': cfg['window'],
         'k': 0.34, 'bg_downsample': cfg['bg'], 'denoise': 'fast' if not cfg['sauvola_only'] else None, 'mask_only': cfg['mask_only'],
         'halftone_page_fraction': cfg['halftone'],
         'parallelism': 'pages sharded over %d GPU(s), no collective' % world,
         'l2': 'inputs (%.2f GB/batch) larger than L2' % (cfg['pages'] * cfg['H'] * cfg['W'] * cfg['C'] / 1e9)}
    c.update(extra)
    return c


# ----------------------------------------------------------------------------- B200 arm
class DeviceRun:
    """One config resident on the GPU: synthetic pages tiled to the batch, a DecomposeBatch, and the step."""

    def __init__(self, cfg, rank, world, pkg, torch):
        self.cfg, self.torch = cfg, torch
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        workers = max(1, min(cfg['distinct'], ncpu // max(world, 1)))
        distinct = make_pages(cfg, rank * cfg['distinct'], cfg['distinct'], workers)
        N, H, W, C = cfg['pages'], cfg['H'], cfg['W'], cfg['C']
        shape = (N, H, W, C) if C == 3 else (N, H, W)
        self.host = torch.empty(shape, dtype=torch.uint8).pin_memory()
        hv = self.host.numpy()
        # halftone pages come first in `distinct`: interleave so that every part of the batch holds its share
        order = np.argsort([(i * 7919) % len(distinct) for i in range(len(distinct))])
        for i in range(N):
            hv[i] = distinct[order[i % len(distinct)]]
        self.eng = pkg.get_engine()
        self.batch = self.eng.make_batch(N, H, W, C, bg_downsample=cfg['bg'], mask_only=cfg['mask_only'])
        self.batch.img.upload(self.host, non_blocking=False)
        torch.cuda.synchronize()
        self.px = N * H * W

    def step(self):
        cfg = self.cfg
        if cfg['sauvola_only']:
            self.eng.sauvola(self.batch.img, self.batch.mask, cfg['window'])
        else:
            self.batch.run(cfg['window'], denoise_mask='fast')

    def timed(self, steps, warmup, barrier):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)


def serialized_kernel_table(run, _lib, steps=3):
    """Per-kernel device ms with nothing overlapping: the same step with one page group on the caller's stream
    (DECOMPOSE_GROUPS = 1), every launch bracketed by a CUDA event pair."""
    torch = run.torch
    saved = _lib.get_tuning('DECOMPOSE_GROUPS')
    _lib.set_tuning('DECOMPOSE_GROUPS', 1)
    cfg = run.cfg
    b = run.eng.make_batch(cfg['pages'], cfg['H'], cfg['W'], cfg['C'], bg_downsample=cfg['bg'], mask_only=cfg['mask_only'])
    b.img = run.batch.img
    try:
        for _ in range(2):
            b.run(cfg['window'], denoise_mask='fast')
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            b.run(cfg['window'], denoise_mask='fast')
        e1.record()
        torch.cuda.synchronize()
        serial_ms = e0.elapsed_time(e1) / steps
        _lib.profile_enable(True)
        for _ in range(steps):
            b.run(cfg['window'], denoise_mask='fast')
        torch.cuda.synchronize()
        ms = {k: v[1] / steps for k, v in _lib.profile_report().items()}       # ms per step (all launches of the kernel)
        _lib.profile_enable(False)
    finally:
        _lib.set_tuning('DECOMPOSE_GROUPS', saved)
    del b
    return ms, serial_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-pipelined', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the short runs of configs 3, 4, 5 in the default line')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 1)
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return run_reference(args, rank, cfg)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')     # NCCL's banner / debug lines go to stderr: stdout carries the one JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    run = DeviceRun(cfg, rank, world, pkg, torch)
    N, H, W, C = cfg['pages'], cfg['H'], cfg['W'], cfg['C']

    # ---- device-resident steps (inputs in HBM, larger than L2: no flush needed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                              # nvidia-smi needs a moment to start: begin before the warm-up
    for _ in range(args.warmup):
        run.step()
    torch.cuda.synchronize()
    t_w = time.time()
    while rank == 0 and sampler.proc and not sampler.rows and time.time() - t_w < 3.0:
        run.step()                                   # extra untimed warm-up until the sampler is alive
        torch.cuda.synchronize()
    barrier()
    launches0 = _lib.lib().b200mrc_launch_count()
    ts0 = sampler.mark()
    dev_ms = max_ranks(run.timed(args.steps, 0, barrier))
    ts1 = sampler.mark()
    launches = _lib.lib().b200mrc_launch_count() - launches0
    clocks = sampler.stop(ts0, ts1) if rank == 0 else None
    px_step = run.px * world
    value = px_step * args.steps / (dev_ms / 1e3) / 1e6
    ms_step = dev_ms / args.steps

    # ---- per-kernel table from a serialised pass of the same step (kernels of different page groups overlap in the
    # timed region above, so their event-pair times there do not add up)
    kernel_ms, serial_ms = ({}, None)
    if not cfg['sauvola_only']:
        kernel_ms, serial_ms = serialized_kernel_table(run, _lib)

    # ---- the same K steps with two batches in flight (one CUDA stream per in-flight batch: a book is a stream of batches):
    # the row-latency-bound sweep of one batch overlaps the issue-bound stages of the next.  Reported next to `value`,
    # which stays the plain one-call-at-a-time figure.
    pipelined = None
    if not args.no_pipelined and not cfg['sauvola_only']:
        batch2 = run.eng.make_batch(N, H, W, C, bg_downsample=cfg['bg'], mask_only=cfg['mask_only'])
        batch2.img = run.batch.img                      # same resident input pages, private outputs / workspace
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        both = [run.batch, batch2]

        def pipelined_steps(k):
            cur = torch.cuda.current_stream()
            for s_ in streams:
                s_.wait_stream(cur)
            for i in range(k):
                with torch.cuda.stream(streams[i % 2]):
                    both[i % 2].run(cfg['window'], denoise_mask='fast')
            for s_ in streams:
                cur.wait_stream(s_)

        pipelined_steps(4)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pipelined_steps(args.steps)
        p1.record()
        barrier()
        pt = max_ranks(p0.elapsed_time(p1))
        pipelined = {'value': px_step * args.steps / (pt / 1e3) / 1e6, 'unit': 'Mpixels/s', 'ms_per_step': pt / args.steps,
                     'batches_in_flight': 2}
        del batch2, both

    # ---- end to end through the public host API: pinned host pages -> results in pinned host memory
    e2e = None
    if not args.no_e2e and not cfg['sauvola_only']:
        from archive_pdf_tools_b200.engine import StreamedDecomposer
        host = run.host
        del run.batch
        torch.cuda.empty_cache()
        e2e_chunk = int(os.environ.get('B200MRC_E2E_CHUNK', '4' if H * W * C > 5e7 else ('8' if C == 3 else '16')))   # 8 RGB pages of 3300x2550: profiles/r2r_e2e_sweep.txt
        e2e_streams = int(os.environ.get('B200MRC_E2E_STREAMS', '2'))
        e2e_buffers = int(os.environ.get('B200MRC_E2E_BUFFERS', '4'))
        last = 'mask' if cfg['mask_only'] else 'bg'

        def measure_e2e(packed, transport='packed'):
            sd = StreamedDecomposer(run.eng, N, H, W, C, chunk=e2e_chunk, bg_downsample=cfg['bg'], buffers=e2e_buffers,
                                    compute_streams=e2e_streams, mask_only=cfg['mask_only'], packed_mask=packed,
                                    mask_transport=transport)
            # a book is a stream of batches: two batches in flight (each with its own pinned result buffers), so the H2D
            # copies and kernels of step i+1 overlap the D2H tail of step i; every step's inputs cross PCIe and every
            # step's results are waited for and read on the host inside the timed region
            outs2 = [sd.alloc_outputs() for _ in range(2)]

            def e2e_run(steps, depth):
                pending = []
                seen = 0
                for i in range(steps):
                    o = outs2[i % 2]
                    pending.append((sd.run_async(host, o, cfg['window'], denoise_mask='fast'), o))
                    if len(pending) >= depth:
                        ev, oo = pending.pop(0)
                        ev.synchronize()                          # the results of that step are in pinned host memory
                        seen += int(oo['mask'][0, 0, 0]) + int(oo[last][-1, -1, -1])
                for ev, oo in pending:
                    ev.synchronize()
                    seen += int(oo['mask'][0, 0, 0]) + int(oo[last][-1, -1, -1])
                return seen

            e2e_run(2, 2)
            e2e_steps = max(3, min(args.steps, 10))
            torch.cuda.synchronize()
            barrier()
            t0 = time.time()
            e2e_run(e2e_steps, 1)                                 # one call at a time (run() semantics)
            torch.cuda.synchronize()
            barrier()
            dt_sync = max_ranks(time.time() - t0)
            t0 = time.time()
            e2e_run(e2e_steps, 2)
            torch.cuda.synchronize()
            barrier()
            dt = max_ranks(time.time() - t0)
            # the platform ceiling of exactly this transfer pattern: every step's H2D and D2H bytes, both directions at
            # once, all ranks concurrently, no kernels (what the host<->device path of this box can deliver at N GPUs)
            s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
            dev_in = torch.empty_like(host, device='cuda')
            # what crosses the bus: with the packed transport the bool plane 'mask' is filled by host workers from 'mask_packed'
            wire = {k: v for k, v in outs2[0].items() if not (k == 'mask' and 'mask_packed' in outs2[0])}
            dev_out = {k: torch.empty_like(v, device='cuda') for k, v in wire.items()}

            def copy_only():
                with torch.cuda.stream(s_h2d):
                    dev_in.copy_(host, non_blocking=True)
                with torch.cuda.stream(s_d2h):
                    for k, v in wire.items():
                        v.copy_(dev_out[k], non_blocking=True)

            copy_only(); torch.cuda.synchronize(); barrier()
            t0 = time.time()
            for _ in range(3):
                copy_only()
            torch.cuda.synchronize()
            barrier()
            dt_copy = max_ranks((time.time() - t0) / 3)
            res = {'value': px_step * e2e_steps / dt / 1e6, 'unit': 'Mpixels/s', 'steps': e2e_steps,
                   'ms_per_step': dt / e2e_steps * 1e3, 'copy_only_ms_per_step': dt_copy * 1e3,
                   'frac_of_copy_ceiling': dt_copy / (dt / e2e_steps),
                   'h2d_bytes_per_step': int(host.numel()) * world,
                   'd2h_bytes_per_step': int(sum(v.numel() for v in wire.values())) * world,
                   'one_call_at_a_time': px_step * e2e_steps / dt_sync / 1e6, 'batches_in_flight': 2, 'packed_mask': packed,
                   'mask_transport': transport}
            sd.close()
            del sd, outs2, dev_in, dev_out
            torch.cuda.empty_cache()
            return res

        e2e = measure_e2e(False)
        e2e['copy_ceiling_note'] = ('copy_only = the same H2D + D2H bytes per step on two streams, no kernels, all ranks at once: the '
                                    'host<->device ceiling of this box at this N; frac = copy_only time / e2e step time')
        e2e['api'] = ('archive_pdf_tools_b200.engine.StreamedDecomposer.run_async: pinned host pages -> 1-D H2D DMA -> device pitching '
                      '(b200mrc_copy2d) -> b200mrc_decompose (%d-page chunks, %d compute streams, %d device buffers) -> device '
                      'unpitching, mask packed to 1 bit per pixel (b200mrc_pack_mask) -> 1-D D2H DMA of mask/fg/bg into pinned host '
                      'buffers -> 3 host worker threads expand the mask rows into the bool plane the reference yields '
                      '(b200mrc_host_unpack_mask); the step is done when the bool plane, fg and bg are in host memory'
                      % (e2e_chunk, e2e_streams, e2e_buffers))
        # the same results with the 1-byte bool plane itself crossing the bus (round-1 behaviour; no host workers)
        e2e['bool_transport_variant'] = measure_e2e(False, 'bool')
        # the same path handing the mask over as PIL mode-'1' rows (b200mrc_pack_mask: what encode_mrc_mask, mrc.py:474-520,
        # builds from the bool array anyway): no host unpack at all.  Reported separately: `value` above returns the
        # reference's own bool plane.
        e2e['packed_mask_variant'] = measure_e2e(True)

    # ---- short runs of the other BASELINE configs (default line only)
    extra = None
    if args.config == 2 and not args.no_extra and world == 1:
        extra = {}
        del run
        torch.cuda.empty_cache()
        for ci in (5, 4, 3):
            c2 = CONFIGS[ci]
            try:
                r2 = DeviceRun(c2, rank, world, pkg, torch)
                ms2 = r2.timed(5, 3, barrier) / 5
                km2, ser2 = serialized_kernel_table(r2, _lib, steps=2)
                peak2 = _peak()[0]
                extra[str(ci)] = {'workload': c2['name'], 'pages_per_gpu': c2['pages'], 'window': c2['window'], 'steps': 5,
                                  'ms_per_step': ms2, 'value': r2.px / (ms2 / 1e3) / 1e6, 'unit': 'Mpixels/s',
                                  'algorithmic_bytes_per_px': c2['bytes_px'],
                                  'roofline_frac': c2['bytes_px'] * r2.px / (ms2 / 1e3) / 1e9 / peak2,
                                  'kernel_ms_serialized': km2, 'ms_per_step_serialized': ser2}
                del r2
            except Exception as ex:                                # an extra config must never take the headline down
                extra[str(ci)] = {'workload': c2['name'], 'error': repr(ex)}
            torch.cuda.empty_cache()

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peak, peak_src = _peak()
    alg = kernel_alg_bytes_px(C, cfg['bg'])
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        pass
    per_kernel = {}
    for k, ms in kernel_ms.items():
        base = k.split('<')[0]
        ab = alg.get(base) if ms >= 0.05 else None           # launches that only skip pages (other blur radii) move no bytes
        per_kernel[k] = {'ms': ms, 'alg_bytes': ab * run_px(cfg) if ab else None,
                         'frac': (ab * run_px(cfg) / (ms / 1e3) / 1e9 / peak) if ab and ms > 0 else None,
                         'dram_bytes': traffic.get(base)}
    dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else None
    step_bytes = cfg['bytes_px'] * run_px(cfg)
    achieved = step_bytes / (ms_step / 1e3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'whole step (one b200mrc_decompose call); dominant kernel: %s' % dom,
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic.get('step_total'), 'peak_source': peak_src,
                'algorithmic_bytes_per_px': cfg['bytes_px'],
                'note': 'achieved = algorithmic bytes of the whole path (inputs once + reference-visible outputs once) / ms_per_step; '
                        'per_kernel: ms per step from a serialised pass of the same step (event-bracketed launches: the bg thumbnail pass, which in the timed step runs beside the sweep as its programmatic dependent, is serialised behind it here, so the rows add up to more than ms_per_step), alg_bytes = the '
                        'kernel\'s own compulsory bytes (DESIGN.md section 3), dram_bytes = ncu dram__bytes per step (profiles/traffic.json)',
                'dominant': dict(per_kernel.get(dom, {}), kernel=dom) if dom else None,
                'per_kernel': per_kernel, 'ms_per_step_serialized': serial_ms}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_sample(cfg)

    line = {
        'metric': METRIC if args.config == 2 else 'MRC decompose Mpixels/sec, ' + cfg['name'],
        'value': value, 'unit': 'Mpixels/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8', 'data': 'synthetic (%d distinct pages per GPU tiled to %d)' % (cfg['distinct'], N),
        'config': bench_config(cfg, world, page_groups=_lib.get_tuning('DECOMPOSE_GROUPS') or 'auto'),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'pipelined': pipelined, 'extra_configs': extra,
    }
    print(json.dumps(line))
    return 0


def run_px(cfg):
    return cfg['pages'] * cfg['H'] * cfg['W']


def _peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    return peak, ('measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s')


if __name__ == '__main__':
    sys.exit(main())

"""Host side of the B200 MRC engine: device buffers (torch is only the allocator / stream /
copy plumbing), the batched device-resident API used for throughput runs, and numpy-in /
numpy-out helpers used by the reference-facing surface in mrc.py and the drop-in shims.

Layout in HBM: a batch is N equally-shaped pages; every plane is "pitched":
uint8 tensor [N, H, pitch] with pitch = round_up(W*C, 16) so rows are 16-byte aligned
(vector loads / TMA-legal); masks are one byte per pixel (numpy bool compatible).
"""
import ctypes as C
import threading

import warnings

import numpy as np
import torch

from . import _lib as L

DENOISE_NONE, DENOISE_FAST, DENOISE_BREGMAN = 'none', 'fast', 'bregman'      # const.py:31-33
BICUBIC, LANCZOS = 0, 1


def _align(v, a=16):
    return (v + a - 1) // a * a


def window_for_dpi(dpi):
    """threshold_image's window rule, internetarchivepdf/mrc.py:70-75."""
    window_size = 51
    if dpi is not None:
        window_size = int(dpi / 4)
        if window_size % 2 == 0:
            window_size += 1
    return window_size


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.B200MrcError('no CUDA device: the b200mrc engine has no CPU fallback')


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Plane:
    """Pitched uint8 device plane batch [N, H, pitch]."""

    def __init__(self, n, h, w, c=1, device=None):
        self.n, self.h, self.w, self.c = n, h, w, c
        self.pitch = _align(w * c)
        self.t = torch.empty((n, h, self.pitch), dtype=torch.uint8, device=device or 'cuda')

    @property
    def ptr(self):
        return C.c_void_p(self.t.data_ptr())

    @property
    def page_stride(self):
        return self.h * self.pitch

    def view(self):
        """[N, H, W*C] strided view without the row padding."""
        return self.t[:, :, : self.w * self.c]

    def upload(self, host, non_blocking=True):
        """host: uint8/bool ndarray or CPU tensor [N,H,W] / [N,H,W,C] (pinned for async copies)."""
        if not isinstance(host, torch.Tensor):
            host = np.ascontiguousarray(host).view(np.uint8)
        if isinstance(host, torch.Tensor):
            ht = host
        elif host.flags.writeable:
            ht = torch.from_numpy(host)
        else:
            # np.asarray(PIL image) is a read-only view of a bytes object: it is only read here, so wrap it as it is
            # instead of copying 25 MB per RGB page to make it writable (torch warns about read-only arrays)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                ht = torch.from_numpy(host)
        ht = ht.reshape(self.n, self.h, self.w * self.c)
        if self.pitch == self.w * self.c:
            self.t.copy_(ht, non_blocking=non_blocking)
        else:
            self.view().copy_(ht.to(self.t.device, non_blocking=non_blocking))
        return self

    def download(self, out=None, non_blocking=False):
        """-> CPU tensor [N,H,W*C] (into `out` if given, e.g. a pinned buffer)."""
        src = self.view()
        if self.pitch != self.w * self.c:
            src = src.contiguous()
        if out is None:
            return src.cpu()
        out.reshape(self.n, self.h, self.w * self.c).copy_(src, non_blocking=non_blocking)
        return out

    def numpy(self, dtype=np.uint8):
        a = self.download().numpy()
        shape = (self.n, self.h, self.w) + ((self.c,) if self.c > 1 else ())
        return a.reshape(shape).view(dtype)


class PageView:
    """One page of a Plane batch as a Plane-like object (n = 1, same memory)."""

    def __init__(self, plane, i):
        self.n, self.h, self.w, self.c, self.pitch = 1, plane.h, plane.w, plane.c, plane.pitch
        self.t = plane.t[i:i + 1]

    ptr = Plane.ptr
    page_stride = Plane.page_stride
    view = Plane.view
    numpy = Plane.numpy
    download = Plane.download


class ThumbnailPlan:
    """PIL Image.thumbnail((int(w/f), int(h/f))) plan (mrc.py:420-434 / 454-468)."""

    def __init__(self, w, h, c, req_w, req_h, reducing_gap=2.0, filter=BICUBIC):
        st = C.c_int(0)
        self.handle = L.lib().b200mrc_thumbnail_plan_create(w, h, c, float(req_w), float(req_h),
                                                            float(reducing_gap or 0.0), filter, C.byref(st))
        if not self.handle:
            L.check(st.value, 'b200mrc_thumbnail_plan_create')
            self.noop = True
            self.out_w, self.out_h = w, h
        else:
            self.noop = False
            ow, oh = C.c_int(0), C.c_int(0)
            L.lib().b200mrc_resample_plan_out_size(self.handle, C.byref(ow), C.byref(oh))
            self.out_w, self.out_h = ow.value, oh.value
        self.c = c

    def workspace_bytes(self, n):
        return 0 if self.noop else L.lib().b200mrc_resample_workspace_bytes(self.handle, n)

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                L.lib().b200mrc_resample_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def downsample_plan(w, h, c, factor, filter=BICUBIC):
    """The reference's guard + thumbnail call: returns (plan_or_None, too_small)."""
    if factor is None:
        return None, False
    wd, hd = int(w / factor), int(h / factor)
    if wd > 0 and hd > 0:
        plan = ThumbnailPlan(w, h, c, wd, hd, 2.0, filter)
        return (None if plan.noop else plan), False
    return None, True


class MrcEngine:
    """One engine per GPU (mrc.get_engine(device)).  Methods taking `Plane`s are asynchronous on the current stream.

    Scratch memory is private to (stage, stream): two streams -- or two host threads on their own streams -- may run
    stages of the same engine concurrently; calls on ONE stream are ordered by the stream.  The optimise workspace is
    never shared with another stage (its mailbox rows must not be overwritten between the FIR pass and the sweep)."""

    def __init__(self, device=None):
        _require_cuda()
        L.lib()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self._ws = {}
        self._lock = threading.Lock()

    # ------------------------------------------------------------------ raw stages (device planes)
    def workspace(self, nbytes, stage='scratch'):
        key = (stage, torch.cuda.current_stream(self.device).cuda_stream)
        with self._lock:
            t = self._ws.get(key)
            if t is None or t.numel() < nbytes:
                t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
                self._ws[key] = t
        return C.c_void_p(t.data_ptr()), t.numel()

    def threshold_mask(self, img, out, window_w, window_h=None, k=0.34, R=128.0, sigma_dev=None, flags=0):
        """create_threshold_mask (mrc.py:300-329): gray -> conditional blur -> Sauvola in one fused pass.
        img: 1- or 3-channel Plane; sigma_dev: float64 device tensor [N] or None (no blur)."""
        need = L.lib().b200mrc_threshold_workspace_bytes(img.w, img.h, img.n)
        wp, wb = self.workspace(need, 'threshold')
        sp = C.c_void_p(sigma_dev.data_ptr()) if sigma_dev is not None else None
        L.check(L.lib().b200mrc_threshold_mask(img.ptr, img.pitch, img.page_stride, img.c, out.ptr, out.pitch, out.page_stride,
                                               img.w, img.h, img.n, window_w, window_h or window_w, float(k), float(R), sp, flags,
                                               wp, wb, _stream_ptr()), 'b200mrc_threshold_mask')

    def sauvola(self, gray, out, window_w, window_h=None, k=0.34, R=128.0, flags=0):
        L.check(L.lib().b200mrc_sauvola(gray.ptr, gray.pitch, gray.page_stride, out.ptr, out.pitch, out.page_stride,
                                        gray.w, gray.h, gray.n, window_w, window_h or window_w, float(k), float(R),
                                        flags, _stream_ptr()), 'b200mrc_sauvola')

    def gray_blur(self, img, gray, sigma_dev=None):
        sp = C.c_void_p(sigma_dev.data_ptr()) if sigma_dev is not None else None
        L.check(L.lib().b200mrc_gray_blur(img.ptr, img.pitch, img.page_stride, img.c, gray.ptr, gray.pitch,
                                          gray.page_stride, img.w, img.h, img.n, sp, _stream_ptr()), 'b200mrc_gray_blur')

    def estimate_noise(self, img):
        """-> float64 device tensor [N] (mrc.py:273-296 on every page)."""
        sigma = torch.empty(img.n, dtype=torch.float64, device=self.device)
        need = L.lib().b200mrc_noise_workspace_bytes(img.w, img.h, img.n)
        wp, wb = self.workspace(need, 'noise')
        L.check(L.lib().b200mrc_estimate_noise(img.ptr, img.pitch, img.page_stride, img.c, img.w, img.h, img.n,
                                               C.c_void_p(sigma.data_ptr()), wp, wb, _stream_ptr()), 'b200mrc_estimate_noise')
        return sigma

    def denoise(self, mask, mincnt=4, n_size=2):
        need = L.lib().b200mrc_denoise_workspace_bytes(mask.w, mask.h, mask.n)
        wp, wb = self.workspace(need, 'denoise')
        L.check(L.lib().b200mrc_denoise(mask.ptr, mask.pitch, mask.page_stride, mask.w, mask.h, mask.n, mincnt, n_size,
                                        wp, wb, _stream_ptr()), 'b200mrc_denoise')

    def optimise(self, mask, img, out_fg=None, n_fg=3, out_bg=None, n_bg=10):
        need = L.lib().b200mrc_optimise_workspace_bytes(img.w, img.h, img.n)
        wp, wb = self.workspace(need, 'optimise')
        f = (out_fg.ptr, out_fg.pitch, out_fg.page_stride) if out_fg is not None else (None, 0, 0)
        b = (out_bg.ptr, out_bg.pitch, out_bg.page_stride) if out_bg is not None else (None, 0, 0)
        L.check(L.lib().b200mrc_optimise(mask.ptr, mask.pitch, mask.page_stride, img.ptr, img.pitch, img.page_stride, img.c,
                                         f[0], f[1], f[2], n_fg, b[0], b[1], b[2], n_bg, img.w, img.h, img.n,
                                         wp, wb, _stream_ptr()), 'b200mrc_optimise')

    def pack_mask(self, mask, invert=False):
        """mask Plane -> uint8 device tensor [N, H, ceil(W/8)]: PIL mode-'1' rows (np.packbits(mask, axis=-1)), the
        form encode_mrc_mask hands to the PNG/JBIG2 encoder (mrc.py:474-520); 8x fewer bytes to bring back."""
        nb = (mask.w + 7) // 8
        out = torch.empty((mask.n, mask.h, nb), dtype=torch.uint8, device=self.device)
        L.check(L.lib().b200mrc_pack_mask(mask.ptr, mask.pitch, mask.page_stride, C.c_void_p(out.data_ptr()), nb, mask.h * nb,
                                          mask.w, mask.h, mask.n, 1 if invert else 0, _stream_ptr()), 'b200mrc_pack_mask')
        return out

    def thumbnail_np(self, arr, req_w, req_h, reducing_gap=2.0, filter=BICUBIC):
        """PIL Image.fromarray(arr).thumbnail((req_w, req_h), resample=filter, reducing_gap=...) on the device;
        arr: uint8 [H,W] or [H,W,3].  Returns a new ndarray (arr itself when Pillow would not resize)."""
        h, w = arr.shape[:2]
        c = 1 if arr.ndim == 2 else arr.shape[2]
        plan = ThumbnailPlan(w, h, c, req_w, req_h, reducing_gap, filter)
        if plan.noop:
            return arr
        src = Plane(1, h, w, c, self.device).upload(arr[None], non_blocking=False)
        dst = Plane(1, plan.out_h, plan.out_w, c, self.device)
        self.resample(plan, src, dst)
        return dst.numpy()[0]

    def resample(self, plan, src, dst):
        need = plan.workspace_bytes(src.n)
        wp, wb = self.workspace(need, 'resample')
        L.check(L.lib().b200mrc_resample(plan.handle, src.ptr, src.pitch, src.page_stride, dst.ptr, dst.pitch,
                                         dst.page_stride, src.n, wp, wb, _stream_ptr()), 'b200mrc_resample')

    # ------------------------------------------------------------------ whole pipeline, device resident
    def make_batch(self, n, h, w, c, bg_downsample=None, fg_downsample=None, mask_only=False):
        """Allocates the input/output planes and the workspace of a batch; returns a DecomposeBatch."""
        return DecomposeBatch(self, n, h, w, c, bg_downsample, fg_downsample, mask_only)

    # ------------------------------------------------------------------ numpy helpers (one-shot)
    def threshold_image_np(self, img, window, k=0.34, R=128.0, raw_inverted=False):
        h, w = img.shape
        src = Plane(1, h, w, 1, self.device).upload(img[None], non_blocking=False)
        dst = Plane(1, h, w, 1, self.device)
        self.sauvola(src, dst, window, window, k, R, L.SAUVOLA_RAW_INVERTED if raw_inverted else 0)
        return dst.numpy(np.bool_)[0]


class DecomposeBatch:
    """Device-resident state of one batch of pages going through create_mrc_hocr_components."""

    def __init__(self, eng, n, h, w, c, bg_downsample=None, fg_downsample=None, mask_only=False):
        self.eng, self.n, self.h, self.w, self.c = eng, n, h, w, c
        dev = eng.device
        self.mask_only = mask_only
        self.errors = set()
        self.img = Plane(n, h, w, c, dev)
        self.mask = Plane(n, h, w, 1, dev)
        self.sigma = torch.zeros(n, dtype=torch.float64, device=dev)
        self.sigma_in = torch.zeros(n, dtype=torch.float64, device=dev)
        self.fg_plan = self.bg_plan = None
        self.fg = self.bg = None
        if not mask_only:
            self.fg_plan, small_f = downsample_plan(w, h, c, fg_downsample)
            self.bg_plan, small_b = downsample_plan(w, h, c, bg_downsample)
            if small_f or small_b:
                self.errors.add('too-small-to-downsample')       # const.py:38, mrc.py:430-431, 464-465
            fw, fh = (self.fg_plan.out_w, self.fg_plan.out_h) if self.fg_plan else (w, h)
            bw, bh = (self.bg_plan.out_w, self.bg_plan.out_h) if self.bg_plan else (w, h)
            self.fg = Plane(n, fh, fw, c, dev)
            self.bg = Plane(n, bh, bw, c, dev)
        self._args = None
        self._ws = None

    def _build_args(self, window, k, R, denoise_fast, use_sigma_in, or_into=False):
        a = L.DecomposeArgs()
        a.img, a.img_pitch, a.img_page_stride, a.channels = self.img.ptr, self.img.pitch, self.img.page_stride, self.c
        a.width, a.height, a.n_pages = self.w, self.h, self.n
        a.window, a.k, a.R = window, k, R
        a.flags = (L.DECOMPOSE_DENOISE_FAST if denoise_fast else 0) | (L.DECOMPOSE_MASK_ONLY if self.mask_only else 0) | \
                  (L.DECOMPOSE_NO_NOISE_EST if use_sigma_in else 0) | (L.DECOMPOSE_OR_INTO_MASK if or_into else 0)
        a.sigma_in = C.c_void_p(self.sigma_in.data_ptr()) if use_sigma_in else None
        a.sigma_out = C.c_void_p(self.sigma.data_ptr())
        a.mask, a.mask_pitch, a.mask_page_stride = self.mask.ptr, self.mask.pitch, self.mask.page_stride
        if not self.mask_only:
            a.fg, a.fg_pitch, a.fg_page_stride = self.fg.ptr, self.fg.pitch, self.fg.page_stride
            a.bg, a.bg_pitch, a.bg_page_stride = self.bg.ptr, self.bg.pitch, self.bg.page_stride
            a.fg_plan = self.fg_plan.handle if self.fg_plan else None
            a.bg_plan = self.bg_plan.handle if self.bg_plan else None
        need = L.lib().b200mrc_decompose_workspace_bytes(C.byref(a))
        if need == 0:
            raise L.B200MrcError('b200mrc_decompose: invalid arguments')
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.eng.device)
        a.workspace, a.workspace_bytes = C.c_void_p(self._ws.data_ptr()), self._ws.numel()
        return a

    def run(self, window, k=0.34, R=128.0, denoise_mask=DENOISE_FAST, sigma=None, or_into_mask=False):
        """Enqueue the whole pipeline (b200mrc_decompose) on the current stream.  `sigma`: optional
        per-page injected sigma_est (host sequence) replacing the on-device estimate.  or_into_mask: self.mask
        already holds the hOCR line masks (mrc.create_hocr_mask) and the page threshold is OR-ed in."""
        if denoise_mask not in (DENOISE_NONE, DENOISE_FAST):
            if denoise_mask == DENOISE_BREGMAN:
                raise NotImplementedError('denoise_bregman is out of scope (SURVEY.md section 2)')
            raise ValueError('Invalid denoise option:', denoise_mask)              # mrc.py:396
        if sigma is not None:
            self.sigma_in.copy_(torch.as_tensor(np.asarray(sigma, np.float64)))
        key = (window, k, R, denoise_mask == DENOISE_FAST, sigma is not None, bool(or_into_mask))
        if self._args is None or self._args[0] != key:
            self._args = (key, self._build_args(window, k, R, denoise_mask == DENOISE_FAST, sigma is not None, bool(or_into_mask)))
        L.check(L.lib().b200mrc_decompose(C.byref(self._args[1]), _stream_ptr()), 'b200mrc_decompose')
        return self

    # stage-by-stage form of run(): the same kernels through the per-stage C-ABI entry points, with a
    # CUDA event recorded between stages so a profile of the step comes from the step itself
    STAGES = ('noise', 'threshold', 'denoise', 'optimise', 'fg_thumbnail', 'bg_thumbnail')

    def run_staged(self, window, k=0.34, R=128.0, denoise_mask=DENOISE_FAST, sigma=None, events=None):
        """events: None or a dict filled with stage -> (start_event, end_event)."""
        eng = self.eng
        if getattr(self, '_staged', None) is None:
            self._staged = True
            self._fg_full = Plane(self.n, self.h, self.w, self.c, eng.device) if (self.fg_plan and not self.mask_only) else None
            self._bg_full = Plane(self.n, self.h, self.w, self.c, eng.device) if (self.bg_plan and not self.mask_only) else None

        def stage(name, fn):
            if events is None:
                return fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            events[name] = (a, b)

        if sigma is None:
            def _noise():
                self.sigma = eng.estimate_noise(self.img)
            stage('noise', _noise)
            sig = self.sigma
        else:
            self.sigma_in.copy_(torch.as_tensor(np.asarray(sigma, np.float64)))
            sig = self.sigma_in
        stage('threshold', lambda: eng.threshold_mask(self.img, self.mask, window, window, k, R, sig))
        if denoise_mask == DENOISE_FAST:
            stage('denoise', lambda: eng.denoise(self.mask, 4, 2))
        if self.mask_only:
            return self
        fgf = self._fg_full if self.fg_plan else self.fg
        bgf = self._bg_full if self.bg_plan else self.bg
        stage('optimise', lambda: eng.optimise(self.mask, self.img, fgf, 3, bgf, 10))
        if self.fg_plan:
            stage('fg_thumbnail', lambda: eng.resample(self.fg_plan, fgf, self.fg))
        if self.bg_plan:
            stage('bg_thumbnail', lambda: eng.resample(self.bg_plan, bgf, self.bg))
        return self


def _copy2d(dst_ptr, dst_pitch, src_ptr, src_pitch, row_bytes, rows, kind, stream):
    L.check(L.lib().b200mrc_copy2d(C.c_void_p(dst_ptr), dst_pitch, C.c_void_p(src_ptr), src_pitch, row_bytes, rows, kind,
                                   C.c_void_p(stream.cuda_stream)), 'b200mrc_copy2d')


class _PendingBatch:
    """What StreamedDecomposer.run_async returns when host workers finish the batch: synchronize() like a CUDA event."""

    def __init__(self, event, futures):
        self.event, self.futures = event, futures

    def synchronize(self):
        self.event.synchronize()
        for f in self.futures:
            f.result()                                             # re-raises a worker's exception

    def query(self):
        return self.event.query() and all(f.done() for f in self.futures)


class StreamedDecomposer:
    """Host-to-host decomposition of a large batch in chunks.  H2D of later chunks, the kernels of up to
    `compute_streams` chunks and D2H of finished chunks run concurrently (copy engines + SMs) over a ring of
    device batches.  Host<->device transfers are flat 1-D DMAs at full PCIe rate into / out of contiguous device
    staging; the (un)pitching is a device-to-device b200mrc_copy2d on the compute stream.  Host tensors should be
    pinned."""

    def __init__(self, eng, n_pages, h, w, c, chunk=8, bg_downsample=None, fg_downsample=None, mask_only=False,
                 buffers=4, compute_streams=2, packed_mask=False, mask_transport='packed', unpack_workers=3):
        self.eng, self.n, self.h, self.w, self.c = eng, n_pages, h, w, c
        # mask_transport='packed' (default): the mask crosses the bus as mode-'1' rows (1 bit per pixel) and worker
        # threads expand it on the host into the bool plane the reference yields (b200mrc_host_unpack_mask) -- same
        # result arrays as 'bool' (the 1-byte plane itself over the bus), 20 % fewer D2H bytes per RGB page: 12.9
        # against 10.8 Gpx/s host to host on one B200, 3 workers at 4.5 GB/s each keep up (profiles/r2r_e2e_sweep.txt).
        # packed_mask=True hands the packed rows themselves to the caller.
        assert mask_transport in ('bool', 'packed')
        self.host_unpack = mask_transport == 'packed' and not packed_mask
        self.pool = None
        if self.host_unpack:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(max_workers=max(1, unpack_workers), thread_name_prefix='b200mrc-unpack')
            packed_mask = True                                     # device side and D2H: exactly the packed hand-off
        self.chunk = max(1, min(chunk, n_pages))
        n_chunks = (n_pages + self.chunk - 1) // self.chunk
        self.nb = max(1, min(buffers, n_chunks))
        self.batches = [eng.make_batch(self.chunk, h, w, c, bg_downsample, fg_downsample, mask_only) for _ in range(self.nb)]
        self.mask_only = mask_only
        # packed_mask: the mask comes back as PIL mode-'1' rows (b200mrc_pack_mask: ceil(W/8) bytes per row, what
        # encode_mrc_mask hands to the JBIG2/PNG encoder, mrc.py:474-520) -- 8x fewer D2H bytes than the bool plane
        self.packed_mask = packed_mask
        self.s_in, self.s_out = torch.cuda.Stream(device=eng.device), torch.cuda.Stream(device=eng.device)
        self.s_cmp = [torch.cuda.Stream(device=eng.device) for _ in range(max(1, min(compute_streams, self.nb)))]
        b = self.batches[0]
        self.out_shapes = dict(mask=(h, (w + 7) // 8) if packed_mask else (h, w), fg=None if mask_only else (b.fg.h, b.fg.w * c), bg=None if mask_only else (b.bg.h, b.bg.w * c))
        self.errors = set(b.errors)
        # contiguous device staging per buffer (planes whose pitch equals the row length need none)
        dev = eng.device

        def stage(plane):
            return None if plane.pitch == plane.w * plane.c else torch.empty((self.chunk, plane.h, plane.w * plane.c), dtype=torch.uint8, device=dev)
        self.st_in = [stage(b_.img) for b_ in self.batches]
        def mask_stage(plane):
            if packed_mask:
                return torch.empty((self.chunk, h, (w + 7) // 8), dtype=torch.uint8, device=dev)
            return stage(plane)
        self.st_out = [dict(mask=mask_stage(b_.mask), fg=None if mask_only else stage(b_.fg), bg=None if mask_only else stage(b_.bg))
                       for b_ in self.batches]

    def close(self):
        """Stop the host unpack workers (idempotent)."""
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def alloc_outputs(self):
        """Pinned host result buffers: mask [N,H,W] u8 (0/1) -- or [N,H,ceil(W/8)] packed rows with packed_mask --,
        fg [N,fh,fw*C], bg [N,bh,bw*C]."""
        o = {'mask': torch.empty((self.n,) + self.out_shapes['mask'], dtype=torch.uint8).pin_memory()}
        if self.host_unpack:
            # the D2H target stays the pinned packed plane; 'mask' is the bool plane the workers fill (CPU writes: pageable)
            o['mask_packed'] = o['mask']
            o['mask'] = torch.empty((self.n, self.h, self.w), dtype=torch.uint8)
        if not self.mask_only:
            o['fg'] = torch.empty((self.n,) + self.out_shapes['fg'], dtype=torch.uint8).pin_memory()
            o['bg'] = torch.empty((self.n,) + self.out_shapes['bg'], dtype=torch.uint8).pin_memory()
        return o

    def run(self, host_pages, out, window, k=0.34, R=128.0, denoise_mask=DENOISE_FAST):
        """host_pages: CPU uint8 tensor [N,H,W(,C)]; out: dict from alloc_outputs().  Returns after
        everything has landed in `out` (one final stream synchronisation)."""
        self.run_async(host_pages, out, window, k, R, denoise_mask).synchronize()
        torch.cuda.current_stream().wait_stream(self.s_out)
        return out

    def run_async(self, host_pages, out, window, k=0.34, R=128.0, denoise_mask=DENOISE_FAST):
        """Enqueue one batch and return the CUDA event that fires when its last result byte is in `out`.
        Calls may follow each other without waiting (a book is a stream of batches): the H2D copies and kernels
        of the next batch then overlap the D2H tail of this one.  Every call in flight needs its own `out`;
        `host_pages` must stay untouched until the event has fired."""
        n, ck, nb = self.n, self.chunk, self.nb
        hp = host_pages.reshape(n, self.h, self.w * self.c)
        assert hp.is_contiguous() and hp.dtype == torch.uint8
        cur = torch.cuda.current_stream()
        for s in [self.s_in, self.s_out] + self.s_cmp:
            s.wait_stream(cur)
        # last users of every device buffer slot, carried over from the previous call
        tail_cmp, tail_out = getattr(self, '_tail_cmp', {}), getattr(self, '_tail_out', {})
        in_done, cmp_done, out_done = {}, {}, {}
        futures = []
        n_chunks = (n + ck - 1) // ck
        names = ('mask',) if self.mask_only else ('mask', 'fg', 'bg')
        for i in range(n_chunks):
            lo, hi = i * ck, min(n, (i + 1) * ck)
            m = hi - lo
            slot = i % nb
            b = self.batches[slot]
            sc = self.s_cmp[slot % len(self.s_cmp)]
            st_in = self.st_in[slot]
            prev_cmp = cmp_done.get(i - nb, tail_cmp.get(slot))
            prev_out = out_done.get(i - nb, tail_out.get(slot))
            with torch.cuda.stream(self.s_in):
                if prev_cmp is not None:
                    self.s_in.wait_event(prev_cmp)                 # input staging / img plane of this slot is free again
                (st_in if st_in is not None else b.img.t)[:m].copy_(hp[lo:hi], non_blocking=True)
                in_done[i] = self.s_in.record_event()
            with torch.cuda.stream(sc):
                sc.wait_event(in_done[i])
                if prev_out is not None:
                    sc.wait_event(prev_out)                        # output staging of this slot was drained
                if st_in is not None:
                    row = self.w * self.c
                    _copy2d(b.img.t.data_ptr(), b.img.pitch, st_in.data_ptr(), row, row, m * self.h, L.COPY_D2D, sc)
                b.run(window, k, R, denoise_mask)                  # b200mrc_decompose on the chunk (a short tail chunk recomputes stale pages: harmless)
                for name in names:
                    plane, st = getattr(b, name), self.st_out[slot][name]
                    if name == 'mask' and self.packed_mask:
                        row_b = (plane.w + 7) // 8
                        L.check(L.lib().b200mrc_pack_mask(plane.ptr, plane.pitch, plane.page_stride, C.c_void_p(st.data_ptr()), row_b,
                                                          plane.h * row_b, plane.w, plane.h, m, 0, C.c_void_p(sc.cuda_stream)),
                                'b200mrc_pack_mask')
                    elif st is not None:
                        row = plane.w * plane.c
                        _copy2d(st.data_ptr(), row, plane.t.data_ptr(), plane.pitch, row, m * plane.h, L.COPY_D2D, sc)
                cmp_done[i] = sc.record_event()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(cmp_done[i])
                for name in names:
                    plane, st = getattr(b, name), self.st_out[slot][name]
                    dst = out['mask_packed'] if (name == 'mask' and self.host_unpack) else out[name]
                    dst[lo:hi].copy_((st if st is not None else plane.t)[:m], non_blocking=True)
                    if name == 'mask' and self.host_unpack:
                        landed = torch.cuda.Event(blocking=True)   # the worker sleeps on it instead of spinning on a core
                        landed.record(self.s_out)
                        futures.append(self.pool.submit(self._unpack, landed, out, lo, hi))
                out_done[i] = self.s_out.record_event()
            tail_cmp[slot], tail_out[slot] = cmp_done[i], out_done[i]
        self._tail_cmp, self._tail_out = tail_cmp, tail_out
        return _PendingBatch(out_done[n_chunks - 1], futures) if self.host_unpack else out_done[n_chunks - 1]

    def _unpack(self, ev, out, lo, hi):
        """Worker thread: wait for the chunk's packed rows to land, expand them into out['mask'][lo:hi] (both calls release the GIL)."""
        ev.synchronize()
        pk, mk = out['mask_packed'], out['mask']
        rb = pk.shape[2]
        L.check(L.lib().b200mrc_host_unpack_mask(C.c_void_p(pk[lo].data_ptr()), rb, self.h * rb, C.c_void_p(mk[lo].data_ptr()), self.w,
                                                 self.h * self.w, self.w, self.h, hi - lo), 'b200mrc_host_unpack_mask')

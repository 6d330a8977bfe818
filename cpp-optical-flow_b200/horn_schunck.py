"""Host-side mirror of the reference interface, on top of the C ABI (include/hs.h).

`hornSchunck` keeps the name, constructor arguments, public fields and method names of
class hornSchunck in /root/reference/HornSchunckOF/hornSchunck.cpp:8-76 so the parity tests read
like the reference's call site (main.cpp:97-98).  `Solver` is the thin RAII wrapper around one
hs_ctx that the bench and the row-slab driver use.  All arithmetic happens in libhs_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import hs_ctypes as H


# cv::Mat depths the reference accepts (convertTo(CV_64FC1), hornSchunck.cpp:23-24) -> hs_frame_dtype
FRAME_DTYPES = {np.dtype(np.uint8): H.FRAME_U8, np.dtype(np.int8): H.FRAME_S8, np.dtype(np.uint16): H.FRAME_U16,
                np.dtype(np.int16): H.FRAME_S16, np.dtype(np.int32): H.FRAME_S32, np.dtype(np.float32): H.FRAME_F32,
                np.dtype(np.float64): H.FRAME_F64}


def _as_u8_image(img: np.ndarray, name: str) -> np.ndarray:
    a = np.asarray(img)
    if a.ndim != 2:
        raise ValueError(f"{name}: expected a single-channel 2-D image, got shape {a.shape} "
                         "(convert colour frames first, main.cpp:11-26)")
    if a.dtype != np.uint8:
        # the reference converts any depth to CV_64F (:23-24); the device path takes 8-bit frames,
        # so accept other dtypes only when the conversion is lossless
        # (a range check first: a negative int8 would wrap to uint8 and back unnoticed)
        if a.size and (a.min() < 0 or a.max() > 255):
            raise ValueError(f"{name}: dtype {a.dtype} holds values outside [0, 255]")
        b = a.astype(np.uint8)
        if not np.array_equal(b.astype(a.dtype), a):
            raise ValueError(f"{name}: dtype {a.dtype} holds values that are not 8-bit integers")
        a = b
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class Solver:
    """One hs_ctx: fixed geometry and parameters, reusable across frame pairs."""

    def __init__(self, width, height, window_size, max_iterations, alpha, batch=1, device=-1,
                 temporal_k=0, flags=0, out_rows=None, stream=None, global_row0=0,
                 devices=None, decomposition=H.DECOMP_BATCH, exchange=H.EXCHANGE_PEER, slab=None,
                 precision=H.PREC_F32, frame_dtype=np.uint8):
        """devices: list of CUDA ordinals -> ONE context over several GPUs (hs_config.num_devices),
        `decomposition` H.DECOMP_BATCH or H.DECOMP_ROW_SLAB.  slab=(rank, world): this process holds
        one row slab of a `height`-row image (one process per GPU); see slab_info / slab_connect."""
        self._lib = H.load_library()
        cfg = H.HsConfig()
        cfg.struct_size = C.sizeof(H.HsConfig)
        cfg.width, cfg.height = int(width), int(height)
        cfg.window_size, cfg.max_iterations = int(window_size), int(max_iterations)
        cfg.alpha = float(alpha)
        cfg.batch, cfg.device, cfg.temporal_k, cfg.flags = int(batch), int(device), int(temporal_k), int(flags)
        if out_rows is not None:
            cfg.out_row_begin, cfg.out_row_end = int(out_rows[0]), int(out_rows[1])
        cfg.stream = stream
        cfg.global_row0 = int(global_row0)
        if devices is not None and len(devices) > 1:
            self._device_ids = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            cfg.num_devices, cfg.device_ids = len(devices), self._device_ids
            cfg.decomposition, cfg.exchange = int(decomposition), int(exchange)
        if slab is not None:
            cfg.slab_rank, cfg.slab_world = int(slab[0]), int(slab[1])
        self.frame_dtype = np.dtype(frame_dtype)
        if self.frame_dtype not in FRAME_DTYPES:
            raise ValueError(f"unsupported frame dtype {self.frame_dtype}")
        cfg.precision, cfg.frame_dtype = int(precision), FRAME_DTYPES[self.frame_dtype]
        self._ctx = C.c_void_p()
        rc = self._lib.hs_create(C.byref(cfg), C.byref(self._ctx))
        if rc != H.HS_OK:
            self._ctx = None
            raise H.HsError(rc, (self._lib.hs_last_error(None) or b"").decode())
        self.width, self.height, self.batch = cfg.width, cfg.height, max(1, cfg.batch)
        self.out_rows = (cfg.out_row_begin, cfg.out_row_end) if out_rows is not None else (0, cfg.height)
        self.flags = cfg.flags
        self.frame_rows = cfg.height + (1 if cfg.flags & H.FLAG_TOP_IS_SEAM else 0) + \
            (1 if cfg.flags & H.FLAG_BOTTOM_IS_SEAM else 0)
        if slab is not None and cfg.slab_world > 1:      # this process holds one band of the image
            info = self.slab_info()
            self.frame_rows = info.frame_end - info.frame_begin
            self.out_rows = (0, info.own_end - info.own_begin)   # download() returns exactly the owned rows

    # -- row slabs across processes ------------------------------------------------------------
    def slab_info(self) -> H.HsSlabInfo:
        info = H.HsSlabInfo()
        self._check(self._lib.hs_get_slab_info(self._ctx, C.byref(info)))
        return info

    def slab_export(self) -> bytes:
        """256 opaque bytes that tell a neighbouring rank where this slab's planes and seam flags live."""
        h = H.HsSlabHandle()
        self._check(self._lib.hs_slab_export(self._ctx, C.byref(h)))
        return bytes(h.bytes)

    def slab_connect(self, up: bytes | None, down: bytes | None):
        def conv(b):
            if b is None:
                return None
            h = H.HsSlabHandle()
            C.memmove(h.bytes, bytes(b), 256)
            return C.byref(h)
        self._check(self._lib.hs_slab_connect(self._ctx, conv(up), conv(down)))

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != H.HS_OK:
            raise H.HsError(rc, (self._lib.hs_last_error(self._ctx) or b"").decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.hs_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _frames(self, prev, nxt):
        """-> (prev, next, row strides, image strides) for (H, W) or (B, H, W) uint8 input."""
        if self.frame_dtype != np.uint8:                  # HS_PREC_F64 contexts take frames of their own depth
            p, n = np.asarray(prev), np.asarray(nxt)
            if p.dtype != self.frame_dtype or n.dtype != self.frame_dtype:
                raise ValueError(f"frames are {p.dtype}/{n.dtype}, context was created for {self.frame_dtype}")
            p = np.ascontiguousarray(p); n = np.ascontiguousarray(n)
            if self.batch == 1 and p.ndim == 2:
                shape, ps, ns, pis, nis = p.shape, p.strides[0], n.strides[0], 0, 0
            else:
                if p.ndim != 3 or p.shape[0] != self.batch:
                    raise ValueError(f"expected {self.batch} frames, got shape {p.shape}")
                shape, ps, ns, pis, nis = p.shape[1:], p.strides[1], n.strides[1], p.strides[0], n.strides[0]
        elif self.batch == 1 and np.asarray(prev).ndim == 2:
            p, n = _as_u8_image(prev, "prev"), _as_u8_image(nxt, "next")
            shape = p.shape
            ps, ns, pis, nis = p.strides[0], n.strides[0], 0, 0
        else:
            p, n = np.asarray(prev), np.asarray(nxt)
            if p.ndim == 3:                               # the same lossless rule as for single frames
                p = np.stack([_as_u8_image(f, "prev") for f in p]); n = np.stack([_as_u8_image(f, "next") for f in n])
            p = np.ascontiguousarray(p, dtype=np.uint8)
            n = np.ascontiguousarray(n, dtype=np.uint8)
            if p.ndim != 3 or p.shape[0] != self.batch:
                raise ValueError(f"expected {self.batch} frames, got shape {p.shape}")
            shape = p.shape[1:]
            ps, ns, pis, nis = p.strides[1], n.strides[1], p.strides[0], n.strides[0]
        if p.shape != n.shape:
            raise ValueError("Image sizes are different. Please provide images of same size.")  # main.cpp:70-73
        if shape != (self.frame_rows, self.width):
            raise ValueError(f"frames are {shape}, context was created for {(self.frame_rows, self.width)}")
        return p, n, ps, ns, pis, nis

    def _outputs(self, dtype):
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("out dtype must be float32 or float64")
        rows = self.out_rows[1] - self.out_rows[0]
        shape = (rows, self.width) if self.batch == 1 else (self.batch, rows, self.width)
        return np.empty(shape, dtype), np.empty(shape, dtype), (H.HS_F64 if dtype == np.float64 else H.HS_F32)

    # -- the reference's two methods ---------------------------------------------------------
    def solve(self, prev, nxt, out_dtype=np.float64):
        """getFlow (hornSchunck.cpp:43-75): host uint8 frames -> host (u, v)."""
        p, n, ps, ns, pis, nis = self._frames(prev, nxt)
        u, v, dt = self._outputs(out_dtype)
        rs = u.strides[-2]
        ims = u.strides[0] if self.batch > 1 else 0
        self._check(self._lib.hs_solve(self._ctx, p.ctypes.data, ps, pis, n.ctypes.data, ns, nis,
                                       u.ctypes.data, rs, ims, v.ctypes.data, rs, ims, dt))
        return u, v

    def solve_async_raw(self, prev_ptr, next_ptr, row_stride, img_stride, u_ptr, v_ptr, out_row_stride, out_img_stride, dt):
        """hs_solve_async on raw (page-locked) host pointers; pair with solve_wait()."""
        self._check(self._lib.hs_solve_async(self._ctx, prev_ptr, row_stride, img_stride, next_ptr, row_stride, img_stride,
                                             u_ptr, out_row_stride, out_img_stride, v_ptr, out_row_stride, out_img_stride, dt))

    def solve_wait(self):
        self._check(self._lib.hs_solve_wait(self._ctx))

    def solve_bgr(self, prev_bgr, next_bgr, out_dtype=np.float64):
        """preprocess() + getFlow (main.cpp:11-26,84,98): 8UC3 BGR frames in, BGR2GRAY on the device."""
        p = np.asarray(prev_bgr)
        n = np.asarray(next_bgr)
        if p.ndim != 3 or p.shape[2] != 3 or p.dtype != np.uint8 or p.shape != n.shape or n.dtype != np.uint8:
            raise ValueError("solve_bgr expects two uint8 H x W x 3 frames of the same size")
        if p.shape[:2] != (self.height, self.width):
            raise ValueError(f"frames are {p.shape[:2]}, context was created for {(self.height, self.width)}")

        def aligned(a):     # the device conversion reads rows as 32-bit words
            a = np.ascontiguousarray(a)
            if a.strides[0] % 4 or a.ctypes.data % 4:
                buf = np.zeros((a.shape[0], (a.shape[1] * 3 + 3) // 4 * 4), np.uint8)
                buf[:, :a.shape[1] * 3] = a.reshape(a.shape[0], -1)
                return buf, buf.strides[0]
            return a, a.strides[0]

        (p, ps), (n, ns) = aligned(p), aligned(n)
        u, v, dt = self._outputs(out_dtype)
        self._check(self._lib.hs_solve_bgr(self._ctx, p.ctypes.data, ps, n.ctypes.data, ns,
                                           u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], dt))
        return u, v

    def gradients(self, prev, nxt, out_dtype=np.float64):
        """getGradients (hornSchunck.cpp:19-41): -> (gradX, gradY, gradT)."""
        p, n, ps, ns, _, _ = self._frames(prev, nxt)
        dtype = np.dtype(out_dtype)
        rows = self.out_rows[1] - self.out_rows[0]
        g = [np.empty((rows, self.width), dtype) for _ in range(3)]
        dt = H.HS_F64 if dtype == np.float64 else H.HS_F32
        self._check(self._lib.hs_gradients(self._ctx, p.ctypes.data, ps, n.ctypes.data, ns,
                                           g[0].ctypes.data, g[1].ctypes.data, g[2].ctypes.data,
                                           g[0].strides[0], dt))
        return tuple(g)

    # -- device-resident pipeline -------------------------------------------------------------
    def upload(self, prev, nxt):
        p, n, ps, ns, pis, nis = self._frames(prev, nxt)
        self._check(self._lib.hs_upload(self._ctx, p.ctypes.data, ps, pis, n.ctypes.data, ns, nis))
        self._keep = (p, n)   # the copy is asynchronous for pinned memory: keep the arrays alive

    def upload_raw(self, prev_ptr, prev_row_stride, prev_img_stride, next_ptr, next_row_stride, next_img_stride):
        self._check(self._lib.hs_upload(self._ctx, prev_ptr, prev_row_stride, prev_img_stride,
                                        next_ptr, next_row_stride, next_img_stride))

    def prepare(self):
        self._check(self._lib.hs_prepare(self._ctx))

    def iterate(self, iterations):
        self._check(self._lib.hs_iterate(self._ctx, int(iterations)))

    def iterate_rows(self, sweeps, row_begin, row_end, flip):
        self._check(self._lib.hs_iterate_rows(self._ctx, int(sweeps), int(row_begin), int(row_end), int(bool(flip))))

    def iterate_until(self, max_sweeps, tolerance, check_every=50):
        """Early exit (extension): -> (sweeps done, residual of the last sweep)."""
        done, res = C.c_int(0), C.c_double(0.0)
        self._check(self._lib.hs_iterate_until(self._ctx, int(max_sweeps), float(tolerance), int(check_every),
                                               C.byref(done), C.byref(res)))
        return done.value, res.value

    def solve_device(self):
        self._check(self._lib.hs_solve_device(self._ctx))

    def sync(self):
        self._check(self._lib.hs_sync(self._ctx))

    def download(self, out_dtype=np.float64):
        u, v, dt = self._outputs(out_dtype)
        rs = u.strides[-2]
        ims = u.strides[0] if self.batch > 1 else 0
        self._check(self._lib.hs_download(self._ctx, u.ctypes.data, rs, ims, v.ctypes.data, rs, ims, dt))
        return u, v

    def download_raw(self, u_ptr, v_ptr, row_stride, img_stride, dt):
        self._check(self._lib.hs_download(self._ctx, u_ptr, row_stride, img_stride, v_ptr, row_stride, img_stride, dt))

    def sample_grid(self, delta=20):
        """u, v of the current device-resident flow at the plot grid (plotFlow.cpp:70-75): rows and
        columns that are multiples of `delta`.  -> two float64 arrays [ceil(H/delta), ceil(W/delta)]."""
        ny, nx = C.c_int(), C.c_int()
        self._check(self._lib.hs_sample_grid(self._ctx, int(delta), None, None, C.byref(ny), C.byref(nx)))
        u = np.empty((ny.value, nx.value), np.float64)
        v = np.empty((ny.value, nx.value), np.float64)
        self._check(self._lib.hs_sample_grid(self._ctx, int(delta), u.ctypes.data, v.ctypes.data, None, None))
        return u, v

    # -- streaming front-end (frame sequences) --------------------------------------------------
    def video_reset(self):
        self._check(self._lib.hs_video_reset(self._ctx))

    def video_push(self, frame, out_dtype=np.float64):
        """Push the next frame of a sequence.  Returns (pair_index, u, v) for the pair that became
        available (one pair behind the frame just pushed), or None while the pipeline fills."""
        f = _as_u8_image(frame, "frame")
        if f.shape != (self.height, self.width):
            raise ValueError(f"frame is {f.shape}, context was created for {(self.height, self.width)}")
        u, v, dt = self._outputs(out_dtype)
        idx = C.c_int(-1)
        self._check(self._lib.hs_video_push(self._ctx, f.ctypes.data, f.strides[0], u.ctypes.data, u.strides[0],
                                            v.ctypes.data, v.strides[0], dt, C.byref(idx)))
        self._keep = f
        return None if idx.value < 0 else (idx.value, u, v)

    def video_flush(self, out_dtype=np.float64):
        u, v, _ = self._outputs(out_dtype)
        idx = C.c_int(-1)
        self._check(self._lib.hs_video_flush(self._ctx, u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0],
                                             C.byref(idx)))
        return None if idx.value < 0 else (idx.value, u, v)

    def device_view(self) -> H.HsDeviceView:
        view = H.HsDeviceView()
        self._check(self._lib.hs_get_device_view(self._ctx, C.byref(view)))
        return view

    def timing(self) -> H.HsTiming:
        t = H.HsTiming()
        self._check(self._lib.hs_get_timing(self._ctx, C.byref(t)))
        return t


class hornSchunck:  # noqa: N801 - the reference's class name (hornSchunck.cpp:8)
    """Drop-in for `hornSchunck hs(windowSize, maxIterations, alpha); hs.getFlow(prev, next, u, v);`
    (main.cpp:97-98).  C++ out-parameters become return values; outputs are float64 H x W arrays,
    freshly allocated, like the CV_64FC1 Mats of the reference."""

    def __init__(self, inpWindowSize, inpMaxIterations, inpAlpha, device=-1, precision="f32"):
        """precision: "f32" = the fast fused path (8-bit frames, within 1e-4 px of the reference);
        "f64" = the reference's own fp64 arithmetic, bit-identical to the fp64 oracle.  Frames that do
        not hold 8-bit integers (the reference takes any depth, :23-24) always use "f64"."""
        self.windowSize = int(inpWindowSize)        # :14
        self.maxIterations = int(inpMaxIterations)  # :15
        self.alpha = float(inpAlpha)                # :16
        self.precision = precision
        self._device = device
        self._solver = None
        self._key = None

    def _ctx_for(self, shape, frame_dtype=np.uint8, f64=False):
        key = (shape, self.windowSize, self.maxIterations, self.alpha, np.dtype(frame_dtype), f64)
        if key != self._key:
            if self._solver is not None:
                self._solver.close()
            self._solver = Solver(shape[1], shape[0], self.windowSize, self.maxIterations, self.alpha,
                                  device=self._device, precision=H.PREC_F64 if f64 else H.PREC_F32,
                                  frame_dtype=frame_dtype)
            self._key = key
        return self._solver

    def _route(self, imagePrev, imageNext):         # noqa: N803
        """-> (solver, prev, next): 8-bit frames on the fp32 path unless "f64" was asked for; anything
        else on the fp64 path in its own depth."""
        p, n = np.asarray(imagePrev), np.asarray(imageNext)
        if p.ndim != 2 or n.ndim != 2:
            raise ValueError("expected single-channel 2-D images (convert colour frames first, main.cpp:11-26)")
        if p.shape != n.shape:
            raise ValueError("Image sizes are different. Please provide images of same size.")  # main.cpp:70-73
        if self.precision != "f64":
            try:
                p8, n8 = _as_u8_image(p, "imagePrev"), _as_u8_image(n, "imageNext")
                return self._ctx_for(p8.shape), p8, n8
            except ValueError:
                pass                                  # not 8-bit integers: the reference's fp64 path
        if p.dtype != n.dtype or p.dtype not in FRAME_DTYPES:
            dt = np.float64 if (p.dtype.kind == "f" or n.dtype.kind == "f" or p.dtype.itemsize > 4) else np.int32
            p, n = p.astype(dt), n.astype(dt)
        return self._ctx_for(p.shape, p.dtype, True), p, n

    def getGradients(self, imagePrev, imageNext):   # noqa: N802,N803
        s, p, n = self._route(imagePrev, imageNext)
        return s.gradients(p, n, np.float64)

    def getFlow(self, imagePrev, imageNext):        # noqa: N802,N803
        s, p, n = self._route(imagePrev, imageNext)
        return s.solve(p, n, np.float64)

    def close(self):
        if self._solver is not None:
            self._solver.close()
            self._solver = None
            self._key = None

"""ctypes binding of include/hs.h - the same symbols the C++ adapter (adapter/hornSchunck.cpp)
calls.  Nothing here computes anything: every call goes to libhs_b200.so (CUDA, sm_100a)."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

HS_OK = 0
HS_F32, HS_F64 = 0, 1
FLAG_TOP_IS_SEAM, FLAG_BOTTOM_IS_SEAM, FLAG_FORCE_GENERIC, FLAG_SINGLE_PHASE, FLAG_TEXTBOOK = 1, 2, 4, 8, 16

PREC_F32, PREC_F64 = 0, 1
FRAME_U8, FRAME_S8, FRAME_U16, FRAME_S16, FRAME_S32, FRAME_F32, FRAME_F64 = range(7)
DECOMP_BATCH, DECOMP_ROW_SLAB = 0, 1
EXCHANGE_PEER, EXCHANGE_NCCL = 0, 1

STATUS_NAMES = {0: "HS_OK", 1: "HS_ERR_INVALID_ARG", 2: "HS_ERR_CUDA", 3: "HS_ERR_OOM",
                4: "HS_ERR_UNSUPPORTED", 5: "HS_ERR_STATE", 6: "HS_ERR_NCCL"}

# every extern "C" symbol include/hs.h declares (tests check the library exports all of them)
SYMBOLS = ["hs_create", "hs_destroy", "hs_solve", "hs_solve_async", "hs_solve_wait", "hs_solve_bgr", "hs_gradients", "hs_upload", "hs_prepare",
           "hs_iterate", "hs_iterate_rows", "hs_iterate_until", "hs_solve_device", "hs_download", "hs_sync", "hs_sample_grid", "hs_get_device_view",
           "hs_video_push", "hs_video_flush", "hs_video_reset", "hs_get_timing", "hs_last_error", "hs_version", "hs_host_alloc", "hs_host_free",
           "hs_get_slab_info", "hs_plan_slab", "hs_slab_export", "hs_slab_connect", "hs_default_temporal_k"]


class HsConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32),
                ("window_size", C.c_int32), ("max_iterations", C.c_int32), ("alpha", C.c_double),
                ("batch", C.c_int32), ("device", C.c_int32), ("temporal_k", C.c_int32),
                ("flags", C.c_uint32), ("out_row_begin", C.c_int32), ("out_row_end", C.c_int32),
                ("stream", C.c_void_p), ("global_row0", C.c_int32),
                ("precision", C.c_int32), ("frame_dtype", C.c_int32),
                ("num_devices", C.c_int32), ("device_ids", C.POINTER(C.c_int32)),
                ("decomposition", C.c_int32), ("exchange", C.c_int32),
                ("slab_world", C.c_int32), ("slab_rank", C.c_int32)]


class HsSlabInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("own_begin", C.c_int32), ("own_end", C.c_int32),
                ("buf_begin", C.c_int32), ("buf_end", C.c_int32), ("frame_begin", C.c_int32), ("frame_end", C.c_int32),
                ("temporal_k", C.c_int32), ("halo_top", C.c_int32), ("halo_bottom", C.c_int32)]


class HsSlabHandle(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 256)]


class HsTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("prepare_ms", C.c_float), ("iterate_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int32),
                ("temporal_k", C.c_int32), ("kernel_id", C.c_int32)]


class HsDeviceView(C.Structure):
    _fields_ = [("prev", C.c_void_p), ("next", C.c_void_p), ("frame_pitch", C.c_size_t),
                ("frame_pair_stride", C.c_size_t), ("frame_rows", C.c_int32), ("frame_row0", C.c_int32),
                ("uv", C.c_void_p), ("reserved", C.c_void_p), ("flow_pitch", C.c_size_t),
                ("flow_pair_stride", C.c_size_t), ("width", C.c_int32), ("height", C.c_int32),
                ("batch", C.c_int32), ("halo_rows_top", C.c_int32), ("halo_rows_bottom", C.c_int32)]


class HsError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


_lib = None


def load_library(path: str | None = None):
    """dlopen libhs_b200.so (building it first if nvcc is here and it is stale).  Raises if the
    library cannot be had: there is deliberately no fallback implementation."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is None and os.environ.get("HS_B200_LIB"):
        path = os.environ["HS_B200_LIB"]          # an A/B build of the same CUDA library (tools/variant_build.sh)
    if path is None:
        path = _build.LIB
        if _build.is_stale() and _build.nvcc_path():
            _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                           " (the CUDA library is the only implementation; there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    lib.hs_create.argtypes = [C.POINTER(HsConfig), C.POINTER(vp)]
    lib.hs_destroy.argtypes = [vp]
    lib.hs_destroy.restype = None
    lib.hs_solve.argtypes = [vp, vp, sz, sz, vp, sz, sz, vp, sz, sz, vp, sz, sz, i32]
    lib.hs_solve_async.argtypes = [vp, vp, sz, sz, vp, sz, sz, vp, sz, sz, vp, sz, sz, i32]
    lib.hs_solve_wait.argtypes = [vp]
    lib.hs_gradients.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp, sz, i32]
    lib.hs_solve_bgr.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, i32]
    lib.hs_upload.argtypes = [vp, vp, sz, sz, vp, sz, sz]
    lib.hs_prepare.argtypes = [vp]
    lib.hs_iterate.argtypes = [vp, i32]
    lib.hs_iterate_rows.argtypes = [vp, i32, i32, i32, i32]
    lib.hs_iterate_until.argtypes = [vp, i32, C.c_double, i32, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.hs_solve_device.argtypes = [vp]
    lib.hs_download.argtypes = [vp, vp, sz, sz, vp, sz, sz, i32]
    lib.hs_sync.argtypes = [vp]
    lib.hs_sample_grid.argtypes = [vp, i32, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.hs_video_push.argtypes = [vp, vp, sz, vp, sz, vp, sz, i32, C.POINTER(C.c_int)]
    lib.hs_video_flush.argtypes = [vp, vp, sz, vp, sz, C.POINTER(C.c_int)]
    lib.hs_video_reset.argtypes = [vp]
    lib.hs_get_device_view.argtypes = [vp, C.POINTER(HsDeviceView)]
    lib.hs_get_timing.argtypes = [vp, C.POINTER(HsTiming)]
    lib.hs_last_error.argtypes = [vp]
    lib.hs_last_error.restype = C.c_char_p
    lib.hs_version.argtypes = []
    lib.hs_host_alloc.argtypes = [C.POINTER(vp), sz]
    lib.hs_host_free.argtypes = [vp]
    lib.hs_default_temporal_k.argtypes = [C.POINTER(HsConfig), i32]
    lib.hs_get_slab_info.argtypes = [vp, C.POINTER(HsSlabInfo)]
    lib.hs_plan_slab.argtypes = [i32, i32, i32, i32, i32, C.POINTER(HsSlabInfo)]
    lib.hs_slab_export.argtypes = [vp, C.POINTER(HsSlabHandle)]
    lib.hs_slab_connect.argtypes = [vp, C.POINTER(HsSlabHandle), C.POINTER(HsSlabHandle)]
    for name in SYMBOLS:
        if name not in ("hs_destroy", "hs_last_error"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib

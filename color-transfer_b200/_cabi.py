"""ctypes binding of libct_b200.so (include/ct_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is
visible, importing the binding or creating a handle raises.
"""

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CT_B200_LIB") or os.path.join(_HERE, "libct_b200.so")  # override: A/B builds only

CT_OK, CT_E_INVALID, CT_E_CUDA, CT_E_NONFINITE, CT_E_NOT_PD, CT_E_SINGULAR, CT_E_UNSUPPORTED, CT_E_NOMEM = (
    0, -1, -2, -3, -4, -5, -6, -7)
CT_F32, CT_F64, CT_U8 = 0, 1, 2
CT_BATCH_U8_AS_F32, CT_BATCH_CLAMP01 = 1, 2   # ct_batch.flags
CT_HWC, CT_CHW = 0, 1
CT_REINHARD, CT_CCS, CT_MKL_MK, CT_MKL_SQRT, CT_MKL_CHOLESKY = 0, 1, 2, 3, 4
CT_MOMENT_DOUBLES = 10
CT_XFORM_DOUBLES = 16
CT_IDT_MAX_BINS = 1024
CT_IDT_KEYS = 6


def lut_doubles(bins):
    return 3 * (3 * ((bins + 2) // 2 * 2) + 4)


class Batch(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("npix", ctypes.c_int64), ("image_stride", ctypes.c_int64),
                ("plane_stride", ctypes.c_int64), ("count", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("layout", ctypes.c_int32), ("flags", ctypes.c_int32)]


class IdtStage(ctypes.Structure):
    _fields_ = [("target", ctypes.POINTER(Batch)), ("reference", ctypes.POINTER(Batch)),
                ("rot", ctypes.c_void_p), ("rot_next", ctypes.c_void_p), ("rot_stride", ctypes.c_int64),
                ("keys", ctypes.c_void_p), ("keys_next", ctypes.c_void_p), ("keys_stride", ctypes.c_int64),
                ("counts", ctypes.c_void_p), ("lut", ctypes.c_void_p), ("status", ctypes.c_void_p),
                ("bins", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class IdtTrace(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_void_p), ("hi", ctypes.c_void_p), ("counts_t", ctypes.c_void_p),
                ("counts_r", ctypes.c_void_p), ("lut", ctypes.c_void_p)]


# name -> (restype, argtypes); also the list tests check against include/ct_b200.h
_P = ctypes.c_void_p
_BP = ctypes.POINTER(Batch)
SIGNATURES = {
    "ct_abi_version": (ctypes.c_int, []),
    "ct_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_P)]),
    "ct_destroy": (None, [_P]),
    "ct_last_error": (ctypes.c_char_p, [_P]),
    "ct_set_stream": (ctypes.c_int, [_P, _P]),
    "ct_synchronize": (ctypes.c_int, [_P]),
    "ct_sm_count": (ctypes.c_int, [_P]),
    "ct_launch_count": (ctypes.c_int64, [_P]),
    "ct_moments": (ctypes.c_int, [_P, _BP, ctypes.c_int, _P]),
    "ct_linear_solve": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, ctypes.c_int, _P, _P]),
    "ct_linear_apply": (ctypes.c_int, [_P, ctypes.c_int, _BP, _P, _BP]),
    "ct_linear_transfer": (ctypes.c_int, [_P, ctypes.c_int, _BP, _BP, _BP, _P, _P]),
    "ct_linear_transfer_host": (ctypes.c_int, [_P, ctypes.c_int, _BP, _BP, _BP]),
    "ct_regrain_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int32, ctypes.c_int32]),
    "ct_regrain": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int32, ctypes.c_int32, _P, ctypes.c_size_t]),
    "ct_acg_transfer_host": (ctypes.c_int, [_P, _BP, _BP, _BP, ctypes.c_int32, ctypes.c_int32, _P, ctypes.c_int32, ctypes.c_int32]),
    "ct_linear_transfer_host_u8": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_int32]),
    "ct_idt_transfer_host_u8": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                               _P, ctypes.c_int32, ctypes.c_int32]),
    "ct_linear_stats_host": (ctypes.c_int, [_P, ctypes.c_int, _BP, _BP, _P, _P]),
    "ct_linear_apply_staged_host": (ctypes.c_int, [_P, ctypes.c_int, _P, _BP]),
    "ct_profile_enable": (ctypes.c_int, [_P, ctypes.c_int]),
    "ct_profile_read": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32]),
    "ct_idt_key_of": (ctypes.c_int64, [ctypes.c_double]),
    "ct_idt_value_of": (ctypes.c_double, [ctypes.c_int64]),
    "ct_idt_keys_init": (ctypes.c_int, [_P, _P, ctypes.c_int64]),
    "ct_idt_ranges": (ctypes.c_int, [_P, _BP, _P, ctypes.c_int64, ctypes.c_int32, _P, ctypes.c_int64, _P]),
    "ct_idt_hist": (ctypes.c_int, [_P, ctypes.POINTER(IdtStage), ctypes.c_int]),
    "ct_idt_lut": (ctypes.c_int, [_P, ctypes.POINTER(IdtStage), ctypes.c_int]),
    "ct_idt_remap": (ctypes.c_int, [_P, ctypes.POINTER(IdtStage), _BP, ctypes.c_int]),
    "ct_idt_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "ct_idt_transfer": (ctypes.c_int, [_P, _BP, _BP, _BP, _P, ctypes.c_int32, ctypes.c_int32, _P, ctypes.c_size_t,
                                       ctypes.POINTER(IdtTrace), _P]),
    "ct_idt_transfer_host": (ctypes.c_int, [_P, _BP, _BP, _BP, _P, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.POINTER(IdtTrace)]),
    "ct_icid": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                               ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]),
    "ct_psnr": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.c_int64, ctypes.POINTER(ctypes.c_double)]),
    "ct_ssim": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                               ctypes.POINTER(ctypes.c_double)]),
    "ct_distort": (ctypes.c_int, [_P, _BP, _P, ctypes.c_int32, _BP]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libct_b200.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python color-transfer_b200/build.py` "
                    "(nvcc, sm_100a). There is no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


class CtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libct_b200 error {code}: {message}")
        self.code = code
        self.message = message


_DTYPES = {np.dtype(np.float32): CT_F32, np.dtype(np.float64): CT_F64}   # uint8 arrays are promoted like the reference's float math; uint8 FRAMES go through batch.*_u8


def batch_from_pointer(ptr, npix, dtype, layout=CT_HWC, count=1, image_stride=0, plane_stride=0):
    return Batch(ctypes.c_void_p(ptr), npix, image_stride, plane_stride, count, _DTYPES[np.dtype(dtype)], layout, 0)


def batch_from_numpy(arr):
    """Describe a host image [H,W,3] (or a batch [B,H,W,3]) without copying when its memory is
    either C-contiguous HWC or the CHW memory of the reference Runner's permuted views
    (ref: methods/__init__.py:21-22); anything else is made contiguous first.  Returns
    (Batch, keepalive array)."""
    a = np.asarray(arr)
    if a.dtype not in _DTYPES:
        a = a.astype(np.float64)
    if a.ndim == 3:
        lead = ()
    elif a.ndim == 4:
        lead = (a.shape[0],)
    else:
        raise ValueError(f"expected [H,W,3] or [B,H,W,3], got shape {a.shape}")
    if a.shape[-1] != 3:
        raise ValueError(f"last dimension must hold 3 channels, got shape {a.shape}")
    count = lead[0] if lead else 1
    npix = int(a.shape[-3] * a.shape[-2])
    if a.flags.c_contiguous:
        layout = CT_HWC
    else:
        chw = np.moveaxis(a, -1, -3)
        if chw.flags.c_contiguous:
            layout = CT_CHW
        else:
            a = np.ascontiguousarray(a)
            layout = CT_HWC
    return Batch(ctypes.c_void_p(a.ctypes.data), npix, 3 * npix, 0, count, _DTYPES[a.dtype], layout, 0), a


class Handle:
    """One libct_b200 context (device + stream + scratch)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.ct_create(int(device), ctypes.byref(h))
        if rc != CT_OK:
            raise CtError(rc, f"ct_create(device={device}) failed: no usable CUDA device (no CPU fallback exists)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.ct_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def check(self, rc):
        if rc != CT_OK:
            raise CtError(rc, self.lib.ct_last_error(self.h).decode("utf-8", "replace"))

    def set_stream(self, stream_ptr):
        self.check(self.lib.ct_set_stream(self.h, ctypes.c_void_p(stream_ptr)))

    def synchronize(self):
        self.check(self.lib.ct_synchronize(self.h))

    PROF_NAMES = {1: "seed", 2: "ranges_target", 3: "ranges_reference", 4: "hist", 5: "remap"}

    def profile(self, on):
        """Start / stop per-launch timing of the fused IDT driver (see ct_profile_enable)."""
        self.check(self.lib.ct_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, max_entries=4096):
        """[(name, ms), ...] of the launches since the last read, in launch order."""
        ids = (ctypes.c_int32 * max_entries)()
        ms = (ctypes.c_float * max_entries)()
        n = self.lib.ct_profile_read(self.h, ids, ms, max_entries)
        if n < 0:
            self.check(n)
        return [(self.PROF_NAMES.get(ids[i], str(ids[i])), float(ms[i])) for i in range(n)]

    @property
    def launches(self):
        return int(self.lib.ct_launch_count(self.h))

    @property
    def sm_count(self):
        return int(self.lib.ct_sm_count(self.h))


_default = {}
_default_lock = threading.Lock()


def default_handle(device=None):
    """Process-wide handle per device (created on first use)."""
    if device is None:
        device = int(os.environ.get("CT_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _default_lock:
        if device not in _default:
            _default[device] = Handle(device)
        return _default[device]

"""Device-resident API: the same transfers on torch CUDA tensors, no host round trip.

PyTorch is only the allocator / stream provider here: tensors are handed to libct_b200.so as raw
device pointers, kernels are ordered on torch's current stream, so ``torch.cuda.Event`` times
them.  ``IdtStages`` exposes the per-stage C-ABI calls (ranges / hist / lut / remap) so that
callers can time each kernel or insert a collective between stages (row-sharded mode).
"""

import ctypes

import torch

from . import _cabi

_TORCH_DTYPE = {torch.float32: _cabi.CT_F32, torch.float64: _cabi.CT_F64, torch.uint8: _cabi.CT_U8}


def _check_images(x, name):
    if not x.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if x.dtype not in _TORCH_DTYPE:
        raise ValueError(f"{name} must be float32, float64 or uint8")
    if x.dim() == 3:
        x = x.unsqueeze(0)
    if x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError(f"{name} must be [H,W,3] or [B,H,W,3]")
    return x


def batch_of(x, flags=0):
    """ct_batch for a [B,H,W,3] tensor that is either contiguous (HWC) or a permuted view of
    contiguous [B,3,H,W] memory (CHW).  ``flags``: _cabi.CT_BATCH_* bits."""
    b, h, w, _ = x.shape
    npix = h * w
    if x.is_contiguous():
        layout, keep = _cabi.CT_HWC, x
    elif x.permute(0, 3, 1, 2).is_contiguous():
        layout, keep = _cabi.CT_CHW, x
    else:
        keep = x.contiguous()
        layout = _cabi.CT_HWC
    return _cabi.Batch(ctypes.c_void_p(keep.data_ptr()), npix, 3 * npix, 0, b, _TORCH_DTYPE[x.dtype], layout, flags), keep


def _handle_for(x, handle):
    h = handle or _cabi.default_handle(x.device.index or 0)
    h.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
    return h


def _in_flags(x, as_float32):
    return _cabi.CT_BATCH_U8_AS_F32 if (x.dtype == torch.uint8 and as_float32) else 0


def _decoded_dtype(x, as_float32):
    """The float dtype the reference would see for this image (a uint8 frame is what its loader decodes)."""
    if x.dtype == torch.uint8:
        return torch.float32 if as_float32 else torch.float64
    return x.dtype


def linear_transfer(method, target, reference, out=None, handle=None, as_float32=True, out_dtype=None, clamp=False):
    """method: _cabi.CT_REINHARD | CT_CCS | CT_MKL_*.  Returns a tensor shaped like target
    (float64, or the target's dtype for Reinhard).  Asynchronous on the current stream.

    uint8 tensors are video frames / ``read_image`` tensors: they are decoded inside the kernels as
    k/255 in float32 (``as_float32``, the reference's dataset loader, ref: utils/data.py:106) or float64
    (``skimage.img_as_float``), and the result defaults to uint8 again (clip + round, img_as_ubyte).
    ``out_dtype`` (torch.uint8 / float32 / float64) picks another result type, converted by the kernel
    that writes it; ``clamp`` applies ``.clamp(0, 1)`` to a float32 result (ref: methods/__init__.py:30)."""
    t = _check_images(target, "target")
    r = _check_images(reference, "reference")
    h = _handle_for(t, handle)
    if out_dtype is None:
        if t.dtype == torch.uint8:
            out_dtype = torch.uint8
        else:
            out_dtype = t.dtype if method == _cabi.CT_REINHARD else torch.float64
    if out is None:
        out = torch.empty(t.shape, dtype=out_dtype, device=t.device)
    o = _check_images(out, "out")
    tb, k1 = batch_of(t, _in_flags(t, as_float32))
    rb, k2 = batch_of(r, _in_flags(r, as_float32))
    ob, k3 = batch_of(o, _cabi.CT_BATCH_CLAMP01 if clamp else 0)
    if k3 is not o:
        raise ValueError("out must be contiguous")
    h.check(h.lib.ct_linear_transfer(h.h, method, tb, rb, ob, None, None))
    return out.view(target.shape) if target.dim() == 3 else out


def idt_transfer(target, reference, rotations, bins=255, n_iter=4, out=None, workspace=None, handle=None,
                 as_float32=True, out_dtype=None, clamp=False):
    """Fused IDT driver (2 + 2*n_iter launches).  rotations: float64 CUDA tensor
    [B, n_iter, 3, 3].  Asynchronous on the current stream.  The result is float64 (the reference's
    dtype) for float images and uint8 for uint8 frames unless ``out`` / ``out_dtype`` say otherwise
    (float32 and uint8 results are written by the last iteration's kernel, n_iter >= 2); see
    ``linear_transfer`` for ``as_float32`` and ``clamp``."""
    t = _check_images(target, "target")
    r = _check_images(reference, "reference")
    h = _handle_for(t, handle)
    b = t.shape[0]
    rot = rotations.reshape(b, n_iter, 9)
    if rot.dtype != torch.float64 or not rot.is_cuda or not rot.is_contiguous():
        raise ValueError("rotations must be a contiguous float64 CUDA tensor [B, n_iter, 3, 3]")
    if out is None:
        if out_dtype is None:
            out_dtype = torch.uint8 if t.dtype == torch.uint8 else torch.float64
        out = torch.empty(t.shape, dtype=out_dtype, device=t.device)
    o = _check_images(out, "out")
    tb, k1 = batch_of(t, _in_flags(t, as_float32))
    rb, k2 = batch_of(r, _in_flags(r, as_float32))
    ob, k3 = batch_of(o, _cabi.CT_BATCH_CLAMP01 if clamp else 0)
    if k3 is not o:
        raise ValueError("out must be a contiguous tensor")
    ws_ptr, ws_bytes = (None, 0)
    if workspace is not None:
        ws_ptr, ws_bytes = ctypes.c_void_p(workspace.data_ptr()), workspace.numel() * workspace.element_size()
    h.check(h.lib.ct_idt_transfer(h.h, tb, rb, ob, ctypes.c_void_p(rot.data_ptr()), bins, n_iter,
                                  ws_ptr, ws_bytes, None, None))
    return out.view(target.shape) if target.dim() == 3 else out


def idt_workspace(npix, count, bins, n_iter, device):
    n = _cabi.load_library().ct_idt_workspace_bytes(npix, count, bins, n_iter)
    return torch.empty(n, dtype=torch.uint8, device=device)


class IdtStages:
    """IDT with every stage a separate call on caller-visible buffers.

    ``between(name, tensor)`` is invoked after the stages that produce a globally reduced
    quantity: ("keys", int64 [B, 6], to be MIN-reduced) once both images have folded their
    projected range of an iteration, and ("counts", int64 [B, 2, 3, bins], to be SUM-reduced)
    after the histograms - the row-sharded driver all-reduces them there.
    ``timer(name)`` may return a context manager to time individual launches.
    """

    def __init__(self, target, reference, rotations, bins=255, n_iter=4, handle=None, as_float32=True):
        self.t = _check_images(target, "target")
        self.r = _check_images(reference, "reference")
        self.as_float32 = as_float32
        self.h = _handle_for(self.t, handle)
        self.bins, self.n_iter = bins, n_iter
        b, hh, ww, _ = self.t.shape
        dev = self.t.device
        self.B, self.npix = b, hh * ww
        self.rot = rotations.reshape(b, n_iter, 9).contiguous()
        self.keys = torch.empty((b, n_iter + 1, _cabi.CT_IDT_KEYS), dtype=torch.int64, device=dev)
        self.counts = torch.zeros((b, 2, 3, bins), dtype=torch.int64, device=dev)
        self.lut_buf = torch.empty((b, _cabi.lut_doubles(bins)), dtype=torch.float64, device=dev)
        self.status = torch.zeros((b,), dtype=torch.int32, device=dev)
        self.plane = (self.npix + 3) // 4 * 4     # 32-byte aligned planes (256-bit state stores)
        self.state = torch.empty((b, 3, self.plane), dtype=torch.float64, device=dev) if n_iter >= 2 else None
        self.out = torch.empty(self.t.shape, dtype=torch.float64, device=dev)
        self.tb, self._k1 = batch_of(self.t, _in_flags(self.t, as_float32))
        self.rb, self._k2 = batch_of(self.r, _in_flags(self.r, as_float32))
        self.ob, _ = batch_of(self.out)
        if self.state is not None:
            self.sb = _cabi.Batch(ctypes.c_void_p(self.state.data_ptr()), self.npix, 3 * self.plane, self.plane, b,
                                  _cabi.CT_F64, _cabi.CT_CHW, 0)

    def _stage(self, it):
        last = it == self.n_iter - 1
        s = _cabi.IdtStage()
        s.target = ctypes.pointer(self.tb if it == 0 else self.sb)
        s.reference = ctypes.pointer(self.rb)
        s.rot = self.rot.data_ptr() + it * 72
        s.rot_next = None if last else self.rot.data_ptr() + (it + 1) * 72
        s.rot_stride = self.n_iter * 9
        s.keys = self.keys.data_ptr() + it * 48
        s.keys_next = None if last else self.keys.data_ptr() + (it + 1) * 48
        s.keys_stride = (self.n_iter + 1) * _cabi.CT_IDT_KEYS
        s.counts = self.counts.data_ptr()
        s.lut = self.lut_buf.data_ptr()
        s.status = self.status.data_ptr()
        s.bins = self.bins
        return s

    # ---- stage protocol (schedule.run_idt_schedule)
    def init(self):
        h = self.h
        self.counts.zero_()
        self.status.zero_()
        h.check(h.lib.ct_idt_keys_init(h.h, ctypes.c_void_p(self.keys.data_ptr()), self.keys.numel()))

    def ranges(self, which):
        h = self.h
        ks, rs = (self.n_iter + 1) * _cabi.CT_IDT_KEYS, self.n_iter * 9
        # the target's range is needed for iteration 0 only; the reference's for every rotation
        h.check(h.lib.ct_idt_ranges(h.h, self.tb if which == "target" else self.rb, ctypes.c_void_p(self.rot.data_ptr()),
                                    rs, 1 if which == "target" else self.n_iter, ctypes.c_void_p(self.keys.data_ptr()), ks,
                                    ctypes.c_void_p(self.status.data_ptr())))

    def hist(self, it, fuse_lut):
        s = self._stage(it)
        self.h.check(self.h.lib.ct_idt_hist(self.h.h, ctypes.byref(s), 1 if fuse_lut else 0))

    def lut(self, it):
        s = self._stage(it)
        self.h.check(self.h.lib.ct_idt_lut(self.h.h, ctypes.byref(s), 0))

    def remap(self, it):
        s = self._stage(it)
        last = it == self.n_iter - 1
        self.h.check(self.h.lib.ct_idt_remap(self.h.h, ctypes.byref(s), self.ob if last else self.sb,
                                             1 if (it == 0 and _decoded_dtype(self.t, self.as_float32) == torch.float32) else 0))

    def result(self):
        return self.out

    def run(self, between=None, timer=None, fuse_lut=None):
        from .schedule import run_idt_schedule
        return run_idt_schedule(self, between, timer, fuse_lut)

    def raise_for_status(self):
        st = self.status.cpu()
        if (st == _cabi.CT_E_NONFINITE).any():
            raise ValueError("supplied range of projected values is not finite")
        if (st != 0).any():
            raise _cabi.CtError(int(st[st != 0][0]), "IDT kernel reported an error")

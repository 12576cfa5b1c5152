"""Stereo video: decode -> transfer -> encode (the step either side of the hot path).

The reference handles video frame by frame on the host: ``cv2.VideoCapture.read`` -> a transfer on
``img_as_float`` frames -> ``img_as_ubyte(...clip(0, 1))`` -> ``cv2.imwrite`` (ref: utils/postprocess.py:78-103,
120-144; the transfer there is ``monge_kantorovitch_color_transfer``, line 138).  Here the decoded uint8
frames go to the device in chunks through pinned memory and the uint8 video API (``batch.*_frames_u8``:
decode / encode fused into the kernels, H2D of chunk k+1 and D2H of chunk k-1 overlapping the kernels of
chunk k inside the C call), while one host thread decodes the next chunk and another encodes the previous
one.  Decoding and encoding stay on the host (OpenCV): they are not part of the statistical transfer path.

Frames are handled in the channel order OpenCV delivers (BGR), exactly like the reference: MKL and IDT do
not depend on the channel order beyond the order itself; pass ``rgb=True`` for Reinhard, whose Lab
conversion does.
"""

import queue
import threading

import numpy as np

from . import batch


def _pinned(shape):
    """uint8 numpy array on pinned host memory (falls back to pageable memory without torch)."""
    try:
        import torch
        return torch.empty(shape, dtype=torch.uint8).pin_memory().numpy()
    except Exception:  # noqa: BLE001
        return np.empty(shape, dtype=np.uint8)


class FrameSource:
    """Iterates (target_frame, reference_frame) uint8 [H,W,3] pairs: two ``cv2.VideoCapture`` streams (paths or
    numbered image patterns OpenCV understands), or any iterable of pairs."""

    def __init__(self, target, reference=None, max_frames=None):
        self.max_frames = max_frames
        if reference is None:
            self._iter = iter(target)
            self.caps = None
        else:
            import cv2
            self.caps = (cv2.VideoCapture(str(target)), cv2.VideoCapture(str(reference)))
            if not all(c.isOpened() for c in self.caps):
                raise IOError(f"Can not open source files: {target}, {reference}")
            self._iter = None

    def __iter__(self):
        n = 0
        while self.max_frames is None or n < self.max_frames:
            if self.caps is None:
                try:
                    t, r = next(self._iter)
                except StopIteration:
                    return
            else:
                ok_t, t = self.caps[0].read()
                ok_r, r = self.caps[1].read()
                if not (ok_t and ok_r):
                    return
            yield t, r
            n += 1

    def close(self):
        if self.caps:
            for c in self.caps:
                c.release()


def transfer_frames(source, sink, method="idt", chunk=8, bins=255, n_iter=4, rgb=False, as_float32=True, handle=None):
    """Pulls frame pairs from ``source`` (FrameSource or iterable), transfers the reference's colours to every target
    frame and hands each uint8 result to ``sink(frame_index, frame)``.  ``method``: "idt" or one of
    ``batch.linear_transfer_frames_u8``'s names ("reinhard", "ccs", "mkl", "mkl_sqrt", "mkl_cholesky").
    IDT rotations are drawn from the global numpy RNG frame by frame, ``n_iter`` per frame - the order a
    sequential loop over the reference function sees.  Returns the number of frames."""
    frames = iter(source)
    # three rotating buffers per stream: one being filled, one queued, one in use - hence queues of depth 1
    todo = queue.Queue(maxsize=1)      # decoded chunk waiting for the device
    done = queue.Queue(maxsize=1)      # result waiting for the encoder
    err = []

    def reader():
        try:
            bufs = None
            slot = 0
            while True:
                n = 0
                for t, r in frames:
                    if bufs is None:
                        bufs = [(_pinned((chunk,) + t.shape), _pinned((chunk,) + r.shape)) for _ in range(3)]
                    tb, rb = bufs[slot]
                    tb[n] = t[..., ::-1] if rgb else t
                    rb[n] = r[..., ::-1] if rgb else r
                    n += 1
                    if n == chunk:
                        break
                if n == 0:
                    break
                todo.put((bufs[slot][0][:n], bufs[slot][1][:n]))
                slot = (slot + 1) % 3
                if n < chunk:
                    break
        except Exception as e:  # noqa: BLE001
            err.append(e)
        finally:
            todo.put(None)

    def writer():
        try:
            index = 0
            while True:
                item = done.get()
                if item is None:
                    break
                for f in item:
                    sink(index, f[..., ::-1] if rgb else f)
                    index += 1
        except Exception as e:  # noqa: BLE001
            err.append(e)
            while done.get() is not None:
                pass

    tr, tw = threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)
    tr.start()
    tw.start()
    total = 0
    outs, k = None, 0
    try:
        while True:
            item = todo.get()
            if item is None or err:
                break
            t8, r8 = item
            if outs is None:
                outs = [_pinned((chunk,) + t8.shape[1:]) for _ in range(3)]
            out = outs[k % 3][:t8.shape[0]]
            k += 1
            if method == "idt":
                batch.idt_frames_u8(t8, r8, bins=bins, n_iter=n_iter, out=out, as_float32=as_float32, handle=handle)
            else:
                batch.linear_transfer_frames_u8(method, t8, r8, out=out, as_float32=as_float32, handle=handle)
            done.put(out)
            total += t8.shape[0]
    finally:
        done.put(None)
        tw.join()
    if err:
        raise err[0]
    return total


def transfer_stereo_video(target_path, reference_path, out_path, method="idt", fourcc="mp4v", fps=None, max_frames=None, **kw):
    """left.mp4 + right.mp4 -> colour-corrected left video (cv2.VideoCapture / cv2.VideoWriter)."""
    import cv2
    src = FrameSource(target_path, reference_path, max_frames)
    if fps is None:
        fps = src.caps[0].get(cv2.CAP_PROP_FPS) or 25.0
    state = {"writer": None}

    def sink(_, frame):
        if state["writer"] is None:
            h, w = frame.shape[:2]
            state["writer"] = cv2.VideoWriter(str(out_path), cv2.VideoWriter_fourcc(*fourcc), fps, (w, h))
            if not state["writer"].isOpened():
                raise IOError(f"Can not open {out_path} for writing (fourcc {fourcc})")
        state["writer"].write(np.ascontiguousarray(frame))

    try:
        return transfer_frames(src, sink, method=method, **kw)
    finally:
        src.close()
        if state["writer"] is not None:
            state["writer"].release()

"""Multi-GPU drivers: one process per GPU, torch.distributed for the plumbing.

Two ways the path partitions:

* frame-parallel - pairs / video frames are independent problems: frame k goes to rank
  k mod world, rotations are pre-drawn in frame order so that a sequential run of the reference
  sees the same matrices, and there is NO data-path collective.
* row-sharded - one oversized pair is split into contiguous row blocks.  The statistics are
  global, and they are exactly additive:
    linear: raw moments {n, S(x-K), S(x-K)(x-K)^T} of both images -> one tiny all-gather,
            combined in rank order on every rank (bit-identical transforms everywhere);
    IDT:    per iteration one int64 MIN all-reduce of the 6 range keys and one int64 SUM
            all-reduce of the 2x3xbins counts (integers: order independent, bit exact).
  The per-pixel work stays local to the shard.

The arithmetic lives behind a small backend protocol so that the collective schedule can be
exercised on CPU (gloo, world_size 2) with a numpy backend supplied by the tests; the product
backend is ``CudaIdtBackend`` / the C-ABI calls and needs a GPU.
"""

import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------ partitions
def frame_partition(n_frames, world, rank):
    """Indices of the frames rank ``rank`` owns: k = rank, rank + world, ..."""
    return list(range(rank, n_frames, world))


def predraw_rotations(n_frames, n_iter, seed=None):
    """[n_frames, n_iter, 3, 3] rotations in the order a sequential loop over the reference's
    ``iterative_distribution_transfer`` would draw them (ref: methods/iterative.py:32)."""
    from .methods.iterative import draw_rotations
    if seed is not None:
        np.random.seed(seed)
    return np.stack([draw_rotations(n_iter) for _ in range(n_frames)])


def row_partition(height, world, rank):
    """Contiguous row block [start, stop) of rank ``rank``; the first ``height % world`` ranks get
    one extra row."""
    base, extra = divmod(height, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


# ------------------------------------------------------------------------------------------ collectives
class Comm:
    """The three exchanges the row-sharded mode needs, on torch tensors (CPU/gloo or CUDA/nccl)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.calls = 0

    def gather_sum_in_rank_order(self, t):
        """all-gather + fixed-order sum: every rank adds the same addends in the same order, so the
        result is bit-identical on all ranks and independent of NCCL's reduction tree."""
        if self.world == 1:
            return t
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        self.calls += 1
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        return acc

    def min_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            self.calls += 1
        return t

    def sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
        return t


# ------------------------------------------------------------------------------------------ linear, row-sharded
def linear_transfer_sharded(method, target_shard, reference_shard, comm=None, handle=None):
    """Closed-form transfer of one pair whose rows are spread over the ranks.  ``*_shard`` are
    this rank's CUDA tensors [rows, W, 3].  Returns this rank's rows of the result."""
    import ctypes

    from . import _cabi, device
    comm = comm or Comm()
    t = device._check_images(target_shard, "target")
    r = device._check_images(reference_shard, "reference")
    h = device._handle_for(t, handle)
    lab = 1 if method == _cabi.CT_REINHARD else 0
    sums = torch.empty((2, _cabi.CT_MOMENT_DOUBLES), dtype=torch.float64, device=t.device)
    tb, _k1 = device.batch_of(t)
    rb, _k2 = device.batch_of(r)
    h.check(h.lib.ct_moments(h.h, tb, lab, ctypes.c_void_p(sums[0].data_ptr())))
    h.check(h.lib.ct_moments(h.h, rb, lab, ctypes.c_void_p(sums[1].data_ptr())))
    sums = comm.gather_sum_in_rank_order(sums)
    xform = torch.empty((_cabi.CT_XFORM_DOUBLES,), dtype=torch.float64, device=t.device)
    status = torch.zeros((1,), dtype=torch.int32, device=t.device)
    h.check(h.lib.ct_linear_solve(h.h, method, ctypes.c_void_p(sums[0].data_ptr()), ctypes.c_void_p(sums[1].data_ptr()),
                                  1, ctypes.c_void_p(xform.data_ptr()), ctypes.c_void_p(status.data_ptr())))
    out = torch.empty(t.shape, dtype=t.dtype if method == _cabi.CT_REINHARD else torch.float64, device=t.device)
    ob, _k3 = device.batch_of(out)
    h.check(h.lib.ct_linear_apply(h.h, method, tb, ctypes.c_void_p(xform.data_ptr()), ob))
    st = int(status.item())
    if st in (_cabi.CT_E_NOT_PD, _cabi.CT_E_SINGULAR):
        raise np.linalg.LinAlgError("Matrix is not positive definite" if st == _cabi.CT_E_NOT_PD else "Singular matrix")
    return out.view(target_shard.shape)


# ------------------------------------------------------------------------------------------ IDT, row-sharded
class CudaIdtBackend:
    """Stage backend on the C ABI (device.IdtStages): buffers are CUDA tensors."""

    def __init__(self, target_shard, reference_shard, rotations, bins, n_iter, handle=None):
        from . import device
        rot = torch.as_tensor(np.ascontiguousarray(rotations, dtype=np.float64)).reshape(1, n_iter, 3, 3)
        self.shape = target_shard.shape
        self.stages = device.IdtStages(target_shard, reference_shard, rot.to(target_shard.device), bins, n_iter, handle)

    def run(self, between):
        return self.stages.run(between=between, fuse_lut=False).view(self.shape)

    def finish(self):
        self.stages.raise_for_status()


def idt_transfer_sharded(target_shard, reference_shard, rotations, bins=255, n_iter=4, comm=None, backend=None,
                         handle=None):
    """IDT of one pair whose rows are spread over the ranks.  Every rank passes the SAME
    ``rotations`` [n_iter,3,3] (draw them on rank 0 and broadcast, or seed identically).
    Collectives: 1 + n_iter MIN all-reduces of 6 int64 keys and n_iter SUM all-reduces of
    6*bins int64 counts.  Returns this rank's rows of the float64 result."""
    comm = comm or Comm()
    backend = backend or CudaIdtBackend(target_shard, reference_shard, rotations, bins, n_iter, handle)

    def between(name, tensor):
        if name == "keys":
            comm.min_(tensor)
        elif name == "counts":
            comm.sum_(tensor)

    out = backend.run(between)
    backend.finish()
    return out

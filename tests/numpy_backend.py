"""TEST-ONLY numpy implementation of the IDT stage protocol (color-transfer_b200/schedule.py) on
oracle arithmetic, so the row-sharded collective schedule can run on CPU under gloo."""

import numpy as np
import torch

from oracle import reference_numpy as oracle

KEY_PLUS_INF = 0x7FF0000000000000


def key_of(x):
    b = np.asarray(x, dtype=np.float64).view(np.int64)
    return np.where(b >= 0, b, b ^ np.int64(0x7FFFFFFFFFFFFFFF))


def value_of(k):
    k = np.asarray(k, dtype=np.int64)
    return np.where(k >= 0, k, k ^ np.int64(0x7FFFFFFFFFFFFFFF)).view(np.float64)


class NumpyIdtBackend:
    def __init__(self, target_shard, reference_shard, rotations, bins, n_iter):
        self.shape = target_shard.shape
        self.state = target_shard.reshape(-1, 3)
        self.ref = reference_shard.reshape(-1, 3)
        self.rot = np.asarray(rotations, dtype=np.float64).reshape(n_iter, 3, 3)
        self.bins, self.n_iter = bins, n_iter
        self.keys = torch.empty((1, n_iter + 1, 6), dtype=torch.int64)
        self.counts = torch.zeros((1, 2, 3, bins), dtype=torch.int64)
        self.trace = []

    # ---- stage protocol
    def init(self):
        self.keys.fill_(KEY_PLUS_INF)
        self.counts.zero_()

    def _fold(self, slot, proj):
        if proj.shape[1] == 0:
            return
        mine = np.concatenate([key_of(proj.min(axis=1)), key_of(-proj.max(axis=1))])
        self.keys[0, slot] = torch.minimum(self.keys[0, slot], torch.from_numpy(mine))

    def ranges(self, which):
        if which == "target":
            self._fold(0, oracle.project(self.rot[0], self.state))
        else:   # the reference is static: every rotation's range up front
            for it in range(self.n_iter):
                self._fold(it, oracle.project(self.rot[it], self.ref))

    def _range(self, it):
        k = self.keys[0, it].numpy()
        return value_of(k[:3]), -value_of(k[3:])

    def hist(self, it, fuse_lut):
        assert not fuse_lut
        lo, hi = self._range(it)
        self.p_t = oracle.project(self.rot[it], self.state)
        p_r = oracle.project(self.rot[it], self.ref)
        self.edges = []
        for j in range(3):
            c_t, c_r, edges = oracle.axis_histograms(self.p_t[j], p_r[j], lo[j], hi[j], self.bins)
            self.counts[0, 0, j] += torch.from_numpy(c_t)
            self.counts[0, 1, j] += torch.from_numpy(c_r)
            self.edges.append(edges)

    def lut(self, it):
        c = self.counts[0].numpy()
        self.luts = [oracle.inverse_cdf_lut(oracle.cdf(c[0, j]), oracle.cdf(c[1, j]), self.edges[j]) for j in range(3)]
        self.trace.append({"counts_t": c[0].copy(), "counts_r": c[1].copy(), "lut": np.stack(self.luts),
                           "lo": self._range(it)[0], "hi": self._range(it)[1]})
        self.counts.zero_()

    def remap(self, it):
        moved = np.empty(self.p_t.shape, dtype=self.state.dtype)
        for j in range(3):
            moved[j] = oracle.remap_axis(self.p_t[j], self.edges[j], self.luts[j], self.bins)
        self.state = oracle.back_rotate(self.rot[it], moved, self.p_t, self.state)
        if it + 1 < self.n_iter:
            self._fold(it + 1, oracle.project(self.rot[it + 1], self.state))

    def result(self):
        return self.state.reshape(self.shape)

    # ---- what sharded.idt_transfer_sharded calls
    def run(self, between):
        from color_transfer_b200.schedule import run_idt_schedule
        return run_idt_schedule(self, between, fuse_lut=False)

    def finish(self):
        pass

"""The stage schedule of one IDT run, shared by the CUDA stage driver (device.IdtStages) and by
the CPU tests' numpy backend: which stage runs when, and where a globally reduced quantity is
handed to ``between`` (the row-sharded driver all-reduces it there).

A stage object provides
    n_iter, keys [B, n_iter+1, 6] int64, counts [B, 2, 3, bins] int64,
    init(), ranges(image) for image in ("target", "reference"),
    hist(it, fuse_lut), lut(it), remap(it).
"""

import contextlib


def run_idt_schedule(st, between=None, timer=None, fuse_lut=None):
    if fuse_lut is None:
        fuse_lut = between is None
    tm = timer or (lambda name: contextlib.nullcontext())
    st.init()
    # K4: the range of iteration 0 needs both images (iterative.py:39-40); the reference is
    # static, so ranges("reference") folds its range under EVERY rotation into keys[:, 0..n_iter-1]
    with tm("ranges_target"):
        st.ranges("target")
    with tm("ranges_reference"):
        st.ranges("reference")
    if between:
        between("keys", st.keys[:, 0])
    for it in range(st.n_iter):
        last = it == st.n_iter - 1
        # K5 (+K6 when fused)
        with tm(f"hist_{it}"):
            st.hist(it, fuse_lut)
        if not fuse_lut:
            if between:
                between("counts", st.counts)
            with tm(f"lut_{it}"):
                st.lut(it)
        # K7: new state; folds the new state's range under the next rotation into keys[:, it + 1]
        with tm(f"remap_{it}"):
            st.remap(it)
        if between and not last:
            between("keys", st.keys[:, it + 1])
    return st.result()

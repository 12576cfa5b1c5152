#!/usr/bin/env python
"""bench.py - stereopair Mpix/s of the B200 colour-transfer path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3] - synthetic 4K VR180 stereo video,
3840x2160 per eye, float32 HWC frames (the reference CLI dtype), iterative distribution
transfer with bins=255, n_iter=4, rotations pre-drawn per frame after np.random.seed(42),
frames sharded frame-parallel over the ranks with no data-path collective (weak scaling:
every rank processes --frames frames per step).  A "step" is one pass of IDT over one batch
of frames.

value      device-resident: frames already in HBM, CUDA-event time on the launching stream.
e2e        the batched host API (color_transfer_b200.batch.idt_frames -> ct_idt_transfer_host):
           pinned host frames in, pinned float64 result out, copies inside the timed region.
roofline   the dominant kernel (largest share of the step), algorithmic bytes / its CUDA-event
           time, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  the numpy oracle port (the reference's own numpy calls) on host cores, bounded
           sample, rank 0 at N=1 only.
--impl reference  the same oracle port with a process pool over frames on all usable cores.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H4K, W4K = 2160, 3840
BINS, N_ITER = 255, 4
METRIC = "stereopair Mpix/s (IDT, 4K stereo video, frame-parallel)"
# algorithmic bytes per target pixel, IDT, float32 frames, n_iter=4 (SURVEY.md 8d / BASELINE.md 4)
IDT_BYTES_PER_PIXEL_F32 = 24 + (24 + 36) + 3 * (36 + 48)


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's "NCCL version ..."
# banner under torchrun), so file descriptor 1 is pointed at stderr for the whole process and the
# result line is written to a private duplicate of the original stdout.
_RESULT_OUT = None


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8, help="frame pairs per rank per step")
    ap.add_argument("--e2e-frames", type=int, default=8)
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--stress", action="store_true", help="i.i.d. uniform frames instead of the smooth field")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--kernels-only", action="store_true", help="profiling aid: only the timed device steps")
    ap.add_argument("--linear-only", action="store_true", help="profiling aid: only the linear-transfer extras")
    ap.add_argument("--workload", default="video4k", choices=["video4k", "rowshard"],
                    help="video4k: configs[3] (default, the bench line); rowshard: configs[4], one 16384x16384 pair "
                         "split by rows over the ranks with NCCL all-reduces of range keys / counts / moments")
    ap.add_argument("--side", type=int, default=16384, help="rowshard: side of the square pair")
    return ap.parse_args()


def ncu_traffic_bytes_per_pixel(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per target pixel of `kernel`, from the committed
    `ncu --set full` capture of this same command (profiles/r01_traffic.json, written by
    tools/ncu_report.py); None if there is no capture for it."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)["dram_bytes_per_pixel"]
        vals = [v for k, vs in t.items() if k.startswith(kernel) for v in vs]
        return sum(vals) / len(vals) if vals else None
    except Exception:  # noqa: BLE001
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.tmp.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_frame(args):
    """One oracle IDT on one synthetic frame pair; runs in a worker process."""
    seed, h, w, rotations, blas_threads = args
    import numpy as np
    if blas_threads:
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(blas_threads)
        except Exception:  # noqa: BLE001
            pass
    from color_transfer_b200_synth import frame_pair  # registered by _register_synth()
    from oracle import reference_numpy as oracle
    t, r = frame_pair(h, w, seed, np.float32)
    t0 = time.perf_counter()
    oracle.iterative_distribution_transfer(t, r, BINS, N_ITER, rotations=rotations)
    return time.perf_counter() - t0


def _register_synth():
    """Load color-transfer_b200/synth.py without importing the package (no CUDA needed)."""
    import importlib.util
    if "color_transfer_b200_synth" in sys.modules:
        return
    spec = importlib.util.spec_from_file_location(
        "color_transfer_b200_synth", os.path.join(ROOT, "color-transfer_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["color_transfer_b200_synth"] = mod
    spec.loader.exec_module(mod)


def usable_workers(bytes_per_worker):
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        n = min(n, max(1, int(psutil.virtual_memory().available * 0.5 // bytes_per_worker)))
    except Exception:  # noqa: BLE001
        pass
    return max(1, min(n, 64))


def cpu_oracle_throughput(h, w, workers, waves, first_seed=2000):
    """Mpix/s of the oracle port: `workers` frames in flight, `waves` rounds."""
    import multiprocessing as mp

    import numpy as np
    import scipy.stats
    _register_synth()
    np.random.seed(42)
    jobs = []
    for k in range(workers * waves):
        rot = np.stack([scipy.stats.special_ortho_group.rvs(3) for _ in range(N_ITER)])
        jobs.append((first_seed + k, h, w, rot, 1 if workers > 1 else 0))
    t0 = time.perf_counter()
    if workers == 1:
        busy = [_cpu_frame(j) for j in jobs]
        wall = sum(busy)          # exclude the synthetic-frame generation
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            busy = pool.map(_cpu_frame, jobs, chunksize=1)
        # frames run concurrently: the wall time of the transfer part is the slowest lane
        lanes = [sum(busy[i::workers]) for i in range(workers)]
        wall = max(lanes)
    total_wall = time.perf_counter() - t0
    mpix = len(jobs) * h * w / 1e6
    return mpix / wall, wall, total_wall


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: quarter-area crops of the 4K frames (IDT cost is linear in pixels)
    h, w = a.height // 2, a.width // 2
    workers = usable_workers(bytes_per_worker=h * w * 3 * 8 * 14)
    values = []
    for _ in range(a.warmup if a.warmup < 2 else 1):
        cpu_oracle_throughput(h, w, workers, 1)
    t_steps = []
    for s in range(a.steps):
        v, wall, _ = cpu_oracle_throughput(h, w, workers, 1, first_seed=2000 + s * workers)
        values.append(v)
        t_steps.append(wall)
    value = sum(values) / len(values)
    sample = (f"{workers} frames in flight per step (one per process), each a {w}x{h} quarter-area frame of the "
              f"4K workload, float32 in, bins={BINS}, n_iter={N_ITER}; numpy oracle port of methods/iterative.py")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * sum(t_steps) / len(t_steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, frames=workers),
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(a, frames):
    return {"workload": "configs[3]: synthetic 4K VR180 stereo video, IDT frame-parallel",
            "frame": [a.height, a.width, 3], "frames_per_rank_per_step": frames, "input_dtype": "float32",
            "bins": BINS, "n_iter": N_ITER, "distribution": "uniform-noise" if a.stress else "smooth-field+noise",
            "parallelism": f"frame-parallel x{a.gpus}, no collective",
            "l2": "working set per frame (99.5 MB/eye in, 199 MB state) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------ our arm
def main():
    a = parse_args()
    _claim_stdout()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import _cabi, batch, device, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    handle = _cabi.default_handle(local)
    H, W, F = a.height, a.width, a.frames
    npix = H * W

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if a.linear_only:
        emit(linear_extras(torch, device, synth, _cabi, handle, dev, measured_peaks()[0]))
        return
    if a.workload == "rowshard":
        run_rowshard(a, torch, dist, device, synth, _cabi, handle, dev, world, rank, barrier, max_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return

    # frames k = rank, rank + world, ...: rotations pre-drawn in frame order after seed 42
    np.random.seed(42)
    all_rot = np.stack([batch.draw_rotations(N_ITER) for _ in range(F * world)])
    my_rot = torch.from_numpy(all_rot[rank::world].copy()).to(dev)
    tgt, ref = synth.frame_pairs_cuda(F, H, W, 2000 + rank, dev, stress=a.stress)
    out = torch.empty((F, H, W, 3), dtype=torch.float64, device=dev)
    ws = device.idt_workspace(npix, F, BINS, N_ITER, dev)

    def step():
        device.idt_transfer(tgt, ref, my_rot, BINS, N_ITER, out=out, workspace=ws, handle=handle)

    for _ in range(max(a.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = handle.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = handle.launches - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / a.steps
    value = world * F * npix / 1e6 / (ms_step / 1e3)
    if a.kernels_only:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "Mpix/s", "ms_per_step": ms_step,
                  "gpu_launches": int(launches), "note": "kernels-only profiling run"})
        return

    # ---- per-kernel breakdown with CUDA events (stage API == the same kernels, unfused LUT off)
    stages = device.IdtStages(tgt, ref, my_rot, BINS, N_ITER, handle=handle)
    times = {}

    class Timer:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()

        def __exit__(self, *exc):
            self.b.record()
            self.b.synchronize()
            times.setdefault(self.name, []).append(self.a.elapsed_time(self.b))

    stages.run(timer=Timer)
    times.clear()
    for _ in range(3):
        stages.run(timer=Timer)
    del stages
    kern = {}
    for name, v in times.items():
        kind = name.split("_")[0]
        kern.setdefault(kind, []).append((name, sum(v) / len(v)))
    share = {k: sum(t for _, t in v) for k, v in kern.items()}
    dominant = max(share, key=share.get)
    # algorithmic bytes per pixel of each launch of the dominant kernel (float32 frames, fp64 state)
    per_launch_bytes = {"hist": lambda it: (12 if it == 0 else 24) + 12, "remap": lambda it: (12 if it == 0 else 24) + 24,
                        "ranges": lambda it: 12}
    achieved = []
    for name, ms in kern[dominant]:
        it = int(name.split("_")[1]) if name.split("_")[1].isdigit() else 0
        achieved.append(per_launch_bytes[dominant](it) * npix * F / (ms / 1e3) / 1e9)
    peak, peak_kind = measured_peaks()
    ach = sum(achieved) / len(achieved)
    kname = {"hist": "hist_kernel", "remap": "remap_kernel", "ranges": "ranges_kernel"}[dominant]
    bpp = ncu_traffic_bytes_per_pixel(kname)
    alg_bytes = sum(per_launch_bytes[dominant](int(n.split("_")[1]) if n.split("_")[1].isdigit() else 0) for n, _ in kern[dominant]) \
        * npix * F / len(kern[dominant])
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": peak, "peak_source": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": None if bpp is None else bpp * npix * F,
                "traffic_source": None if bpp is None else "profiles/r01_traffic.json: ncu dram bytes per pixel of the steady-state launches x pixels per launch",
                "kernel_ms": {k: round(sum(t for _, t in v), 4) for k, v in kern.items()},
                "step_algorithmic_GBps": IDT_BYTES_PER_PIXEL_F32 * npix * F / (ms_step / 1e3) / 1e9,
                "step_frac": IDT_BYTES_PER_PIXEL_F32 * npix * F / (ms_step / 1e3) / 1e9 / peak}

    # ---- end to end through the batched host API, pinned buffers, copies inside the timed region
    Fe = min(a.e2e_frames, F)
    ht = batch.pinned_empty((Fe, H, W, 3), np.float32)
    hr = batch.pinned_empty((Fe, H, W, 3), np.float32)
    ho = batch.pinned_empty((Fe, H, W, 3), np.float64)
    ht[...] = tgt[:Fe].cpu().numpy()
    hr[...] = ref[:Fe].cpu().numpy()
    rot_e2e = all_rot[rank::world][:Fe]
    handle.set_stream(0)
    batch.idt_frames(ht, hr, BINS, N_ITER, rotations=rot_e2e, out=ho, handle=handle)
    e2e_steps = max(2, min(a.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.idt_frames(ht, hr, BINS, N_ITER, rotations=rot_e2e, out=ho, handle=handle)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e = {"value": world * Fe * npix / 1e6 / e2e_s, "unit": "Mpix/s", "h2d_bytes_per_step": int(2 * Fe * npix * 12),
           "d2h_bytes_per_step": int(Fe * npix * 24), "frames_per_step": Fe, "ms_per_step": e2e_s * 1e3,
           "api": "color_transfer_b200.batch.idt_frames -> ct_idt_transfer_host"}
    # the device-resident result and the host-API result must agree (same kernels)
    same = bool(np.array_equal(ho[0], out[0].cpu().numpy()))

    # ---- the same frames as uint8 video frames: 3 B/px over PCIe each way
    h8t = batch.pinned_empty((Fe, H, W, 3), np.uint8)
    h8r = batch.pinned_empty((Fe, H, W, 3), np.uint8)
    h8o = batch.pinned_empty((Fe, H, W, 3), np.uint8)
    h8t[...] = np.rint(ht * 255).astype(np.uint8)
    h8r[...] = np.rint(hr * 255).astype(np.uint8)
    batch.idt_frames_u8(h8t, h8r, BINS, N_ITER, rotations=rot_e2e, out=h8o, handle=handle)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.idt_frames_u8(h8t, h8r, BINS, N_ITER, rotations=rot_e2e, out=h8o, handle=handle)
    barrier()
    u8_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    want8 = np.rint(np.clip(ho[0], 0, 1) * 255).astype(np.uint8)
    e2e_u8 = {"value": world * Fe * npix / 1e6 / u8_s, "unit": "Mpix/s", "h2d_bytes_per_step": int(2 * Fe * npix * 3),
              "d2h_bytes_per_step": int(Fe * npix * 3), "frames_per_step": Fe, "ms_per_step": u8_s * 1e3,
              "api": "color_transfer_b200.batch.idt_frames_u8 -> ct_idt_transfer_host_u8 (uint8 frames in and out)",
              "matches_float_path_bytes": float(np.mean(h8o[0] == want8))}

    extras = {}
    if not a.no_extras and rank == 0:
        extras = linear_extras(torch, device, synth, _cabi, handle, dev, peak)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, wall, _ = cpu_oracle_throughput(H, W, 1, 2)      # two full frame pairs: ~12 s of CPU work
        cpu = {"value": v, "unit": "Mpix/s", "cores": 1, "kind": "port",
               "sample": f"two {W}x{H} frame pairs of the workload (of the {F} per step), float32 in, bins={BINS}, "
                         f"n_iter={N_ITER}, single process ({wall:.1f} s); numpy oracle port of methods/iterative.py"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, F),
                "clocks": clocks, "e2e": e2e, "e2e_u8": e2e_u8, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "host_api_matches_device_api": same, "extras": extras}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_rowshard(a, torch, dist, device, synth, _cabi, handle, dev, world, rank, barrier, max_over_ranks):
    """configs[4]: one side x side float32 pair, contiguous row blocks per rank (strong scaling).
    IDT: 1 + n_iter MIN all-reduces of 6 int64 keys, n_iter SUM all-reduces of 6*bins int64 counts;
    MKL / Reinhard: one all-gather of 2x10 raw moments.  Device-resident shards, CUDA events,
    max over ranks."""
    import numpy as np
    from color_transfer_b200 import sharded
    side = a.side
    r0, r1 = sharded.row_partition(side, world, rank)
    rows = r1 - r0
    # every rank generates only its rows (seeded per row block so the pair is the same for any world size)
    tgt, ref = synth.frame_pairs_cuda(1, rows, side, 3000 + r0, dev)
    tgt, ref = tgt[0], ref[0]
    np.random.seed(42)
    rot = sharded.predraw_rotations(1, N_ITER)[0]
    comm = sharded.Comm()
    results = {}
    peak, _ = measured_peaks()
    npix = side * side

    def timed(fn, reps):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / reps

    backend = sharded.CudaIdtBackend(tgt, ref, rot, BINS, N_ITER, handle)

    def idt():
        def between(name, tensor):
            comm.min_(tensor) if name == "keys" else comm.sum_(tensor)
        backend.run(between)

    ms = timed(idt, max(2, min(a.steps, 5)))
    results["idt"] = {"ms_per_pair": ms, "Mpix/s": npix / 1e6 / (ms / 1e3),
                      "frac_of_hbm_aggregate": IDT_BYTES_PER_PIXEL_F32 * npix / (ms / 1e3) / 1e9 / (peak * world),
                      "collectives_per_pair": 1 + 2 * N_ITER - 1 + 1}
    for name, code, bpp in (("mkl", _cabi.CT_MKL_MK, 60), ("reinhard", _cabi.CT_REINHARD, 48)):
        ms = timed(lambda: sharded.linear_transfer_sharded(code, tgt, ref, comm=comm, handle=handle), max(2, min(a.steps, 5)))
        results[name] = {"ms_per_pair": ms, "Mpix/s": npix / 1e6 / (ms / 1e3),
                         "frac_of_hbm_aggregate": bpp * npix / (ms / 1e3) / 1e9 / (peak * world), "collectives_per_pair": 1}
    if rank == 0:
        emit({"metric": "stereopair Mpix/s (one %dx%d pair, row-sharded)" % (side, side), "unit": "Mpix/s",
              "n_gpus": world, "scaling": "strong", "value": results["idt"]["Mpix/s"], "data": "synthetic",
              "config": {"workload": "configs[4]: single %dx%d float32 stereopair row-sharded with NCCL all-reduces" % (side, side),
                         "rows_per_rank": rows, "bins": BINS, "n_iter": N_ITER}, "results": results})


def linear_extras(torch, device, synth, _cabi, handle, dev, peak):
    """Secondary numbers for the linear transfers (configs[2] shape: 960x540 float32 pairs)."""
    out = {}
    B, H, W = 64, 540, 960
    tgt, ref = synth.frame_pairs_cuda(B, H, W, 1000, dev)
    for name, method, bpp in (("reinhard_f32", _cabi.CT_REINHARD, 48), ("mkl_f32_to_f64", _cabi.CT_MKL_MK, 60)):
        dst = torch.empty((B, H, W, 3), dtype=torch.float32 if method == _cabi.CT_REINHARD else torch.float64, device=dev)
        for _ in range(3):
            device.linear_transfer(method, tgt, ref, out=dst, handle=handle)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 10
        for _ in range(reps):
            device.linear_transfer(method, tgt, ref, out=dst, handle=handle)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbps = bpp * B * H * W / (ms / 1e3) / 1e9
        out[name] = {"Mpix/s": B * H * W / 1e6 / (ms / 1e3), "ms_per_batch": ms, "pairs": B, "shape": [H, W, 3],
                     "algorithmic_GBps": gbps, "frac_of_hbm": gbps / peak,
                     "note": "batch working set 1.2-2.0 GB, larger than L2"}
    # configs[2] as named: the whole batch of 1035 pairs of 960x540 float32 in one call, primary
    # (smooth field) and stress (i.i.d. uniform) distributions
    del tgt, ref, dst
    B2 = 1035
    full = {}
    for dist_name, stress in (("smooth-field+noise", False), ("uniform-noise", True)):
        tgt, ref = synth.frame_pairs_cuda(B2, H, W, 1000, dev, stress=stress)
        for name, method, bpp in (("reinhard_f32", _cabi.CT_REINHARD, 48), ("mkl_f32_to_f64", _cabi.CT_MKL_MK, 60)):
            if stress and method != _cabi.CT_REINHARD:
                continue
            dst = torch.empty((B2, H, W, 3), dtype=torch.float32 if method == _cabi.CT_REINHARD else torch.float64, device=dev)
            device.linear_transfer(method, tgt, ref, out=dst, handle=handle)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                device.linear_transfer(method, tgt, ref, out=dst, handle=handle)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            gbps = bpp * B2 * H * W / (ms / 1e3) / 1e9
            full[f"{name}/{dist_name}"] = {"Mpix/s": B2 * H * W / 1e6 / (ms / 1e3), "ms_per_batch": ms,
                                           "algorithmic_GBps": gbps, "frac_of_hbm": gbps / peak}
            del dst
        del tgt, ref
    out["config2_1035_pairs_960x540"] = full
    # configs[0] / configs[1] shape: one 1080x860 float64 pair (L2-resident, launch-latency regime)
    import numpy as np
    from color_transfer_b200 import batch
    t64, r64 = synth.frame_pairs_cuda(1, 860, 1080, 964, dev, dtype=torch.float64)
    np.random.seed(42)
    rot = torch.from_numpy(batch.draw_rotations(N_ITER)[None]).to(dev)
    single = {"reinhard": lambda: device.linear_transfer(_cabi.CT_REINHARD, t64, r64, handle=handle),
              "mkl": lambda: device.linear_transfer(_cabi.CT_MKL_MK, t64, r64, handle=handle),
              "idt": lambda: device.idt_transfer(t64, r64, rot, BINS, N_ITER, handle=handle)}
    lat = {}
    for name, fn in single.items():
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        lat[name] = {"us_per_pair": us, "Mpix/s": 860 * 1080 / us}
    out["single_pair_1080x860_f64"] = dict(lat, note="device-resident, back-to-back calls, working set fits L2")
    out["dropin_pair0964"] = dropin_pair0964()
    return out


def dropin_pair0964():
    """configs[0] / configs[1]: the reference's own stereo pair through the reference-facing numpy
    functions (pageable float64 arrays in, new array out, H2D/D2H inside the call), wall clock per
    call, next to the CPU oracle on the same arrays."""
    import numpy as np
    try:
        from PIL import Image
        left = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "0964_L.png")).convert("RGB")) / 255.0
        right = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "0964_R.png")).convert("RGB")) / 255.0
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)}
    import methods.iterative
    import methods.linear
    from oracle import reference_numpy as oracle
    res = {"shape": list(left.shape), "dtype": "float64"}
    cases = (("color_transfer_between_images", methods.linear.color_transfer_between_images, oracle.color_transfer_between_images),
             ("monge_kantorovitch_color_transfer", methods.linear.monge_kantorovitch_color_transfer, oracle.monge_kantorovitch_color_transfer),
             ("iterative_distribution_transfer", methods.iterative.iterative_distribution_transfer, oracle.iterative_distribution_transfer),
             ("automated_color_grading", methods.iterative.automated_color_grading, oracle.automated_color_grading))
    for name, ours, ref in cases:
        time.sleep(0.5)     # let the BLAS worker threads of the previous case's oracle call stop spinning
        np.random.seed(42)
        ours(left, right)
        t0 = time.perf_counter()
        for _ in range(5):
            np.random.seed(42)
            got = ours(left, right)
        t_ours = (time.perf_counter() - t0) / 5
        np.random.seed(42)
        t0 = time.perf_counter()
        want = ref(left, right)
        t_ref = time.perf_counter() - t0
        res[name] = {"ms_per_call": t_ours * 1e3, "cpu_oracle_ms": t_ref * 1e3, "speedup": t_ref / t_ours,
                     "max_abs_diff": float(np.max(np.abs(got - want)))}
    return res


if __name__ == "__main__":
    main()

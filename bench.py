#!/usr/bin/env python
"""bench.py - stereopair Mpix/s of the B200 colour-transfer path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3] - synthetic 4K VR180 stereo video,
3840x2160 per eye, iterative distribution transfer with bins=255, n_iter=4, rotations pre-drawn
per frame after np.random.seed(42), frames sharded frame-parallel over the ranks with no
data-path collective (weak scaling: every rank processes the same number of frames per step).
A "step" is `--passes` passes of IDT over the rank's `--frames` resident frame pairs
(8 x 12 = 96 frame pairs per rank and step by default, ~55 ms, so that K = 20 steps time > 1 s).

value      device-resident float32 frames (the reference CLI dtype), CUDA-event time on the launching
           stream, max over ranks.
e2e        the headline: the video API (color_transfer_b200.batch.idt_frames_u8 ->
           ct_idt_transfer_host_u8): pinned uint8 frames in, pinned uint8 frames out, the kernels
           decode / encode as they read / write, copies inside the timed region.
e2e_float  the same through the float API (float32 frames in, the reference's float64 result out).
host_dma   what the box's host<->device DMA gives ALL ranks together (pure pinned copies, both
           directions at once): the ceiling of any end-to-end number at N GPUs.
roofline   the dominant kernel (largest share of the step): algorithmic bytes / its CUDA-event time
           (per-launch events inside the fused driver, ct_profile_*), against MEASURED_PEAKS.json.
linear     configs[2]: 1035 pairs of 960x540 float32, Reinhard / MKL / CCS, frame-parallel, both
           distributions, per-rank device-resident batches (weak scaling like the main metric).
rowshard   (N > 1) configs[4]: one 16384x16384 float32 pair, rows sharded over the ranks with NCCL
           all-reduces of range keys / counts / moments, plus a parity check of the sharded result
           against the unsharded one on a 4096x4096 pair of the same generator.
cpu_baseline  the numpy oracle port on ONE host core, bounded sample, rank 0 at N=1 only.
--impl reference  the same oracle port with a process pool over frames on all usable cores, on the
           same 4K uint8 frames (decode -> IDT -> encode).
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
# ... uint8 frames in and out (3 B/px images, fp64 state): ranges 6, hist 6 + 3*27, remap 27 + 2*48 + 27
IDT_BYTES_PER_PIXEL_U8 = 6 + (6 + 3 * 27) + (27 + 2 * 48 + 27)


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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8, help="resident frame pairs per rank")
    ap.add_argument("--passes", type=int, default=12, help="passes over the resident frames per step")
    ap.add_argument("--e2e-frames", type=int, default=32, help="frame pairs per end-to-end step (one API call)")
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--stress", action="store_true", help="i.i.d. uniform frames instead of the smooth field")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the linear / rowshard / single-pair sections")
    ap.add_argument("--kernels-only", action="store_true", help="profiling aid: only the timed device steps")
    ap.add_argument("--linear-only", action="store_true", help="profiling aid: only the linear section")
    ap.add_argument("--linear-pairs", type=int, default=1035)
    ap.add_argument("--rowshard-side", type=int, default=16384, help="side of the square row-sharded pair")
    ap.add_argument("--ref-workers", type=int, default=0, help="reference arm: processes (0 = all usable cores)")
    return ap.parse_args()


def ncu_traffic_bytes_per_pixel(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per target pixel of `kernel`, from the committed
    `ncu --set full` capture of this same command (profiles/r02_traffic.json, else r01; written by
    tools/ncu_report.py); None if there is no capture for it."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)["dram_bytes_per_pixel"]
            vals = [v for k, vs in t.items() if k.startswith(kernel) for v in vs]
            if vals:
                return sum(vals) / len(vals), "profiles/" + name
        except Exception:  # noqa: BLE001
            continue
    return None, None


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
    """One frame pair of the video path on the CPU: uint8 frames -> k/255 float32 (the reference loader,
    utils/data.py:106) -> oracle IDT -> clip + round to uint8 (img_as_ubyte).  Runs in a worker process."""
    seed, h, w, rotations, blas_threads = args
    import numpy as np
    if blas_threads:
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(blas_threads)
        except Exception:  # noqa: BLE001
            pass
    from color_transfer_b200_synth import frame_pair_u8  # registered by _register_synth()
    from oracle import reference_numpy as oracle
    t8, r8 = frame_pair_u8(h, w, seed)
    t0 = time.perf_counter()
    t = t8 / np.float32(255)
    r = r8 / np.float32(255)
    out = oracle.iterative_distribution_transfer(t, r, BINS, N_ITER, rotations=rotations)
    np.rint(np.clip(out, 0, 1) * 255).astype(np.uint8)
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
        n = min(n, max(1, int(psutil.virtual_memory().available * 0.6 // bytes_per_worker)))
    except Exception:  # noqa: BLE001
        pass
    return max(1, min(n, 128))


class CpuPool:
    """A pool of worker processes that stays up for the whole reference run (steps time only the frames)."""

    def __init__(self, workers):
        import multiprocessing as mp
        _register_synth()
        self.workers = workers
        self.pool = mp.get_context("fork").Pool(workers) if workers > 1 else None

    def step(self, h, w, first_seed, rotations):
        """One frame pair per worker, all in flight together; returns (Mpix/s, wall seconds of the step)."""
        jobs = [(first_seed + k, h, w, rotations[k], 1 if self.workers > 1 else 0) for k in range(self.workers)]
        t0 = time.perf_counter()
        if self.pool is None:
            busy = [_cpu_frame(j) for j in jobs]
            wall = sum(busy)          # exclude the synthetic-frame generation
        else:
            busy = self.pool.map(_cpu_frame, jobs, chunksize=1)
            wall = max(busy)          # frames run concurrently: the step lasts as long as its slowest frame
        total = time.perf_counter() - t0
        return self.workers * h * w / 1e6 / wall, wall, total

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def draw_rotations_np(n_frames):
    import numpy as np
    import scipy.stats
    return [np.stack([scipy.stats.special_ortho_group.rvs(3) for _ in range(N_ITER)]) for _ in range(n_frames)]


def run_reference_arm(a):
    """The reference's own CPU path (the numpy oracle port of methods/iterative.py; the reference itself is
    not importable on the box) on the SAME workload: 3840x2160 uint8 frame pairs, bins / n_iter / rotations
    as in our arm.  A step is a bounded sample of it: one frame pair per host core, all in flight."""
    import numpy as np
    if int(os.environ.get("RANK", "0")) != 0:
        return
    h, w = a.height, a.width
    workers = a.ref_workers or usable_workers(bytes_per_worker=h * w * 3 * 8 * 16)
    pool = CpuPool(workers)
    np.random.seed(42)
    warm = max(a.warmup, 0)
    values, walls = [], []
    for s in range(warm + a.steps):
        rot = draw_rotations_np(workers)
        v, wall, _ = pool.step(h, w, 2000 + s * workers, rot)
        if s >= warm:
            values.append(v)
            walls.append(wall)
    pool.close()
    value = sum(values) / len(values)
    sample = (f"each step = {workers} frame pairs of the workload ({w}x{h} uint8 per eye) in flight together, one per process "
              f"(BLAS pinned to 1 thread per process): decode k/255 float32 -> numpy oracle port of methods/iterative.py "
              f"(bins={BINS}, n_iter={N_ITER}) -> clip + round to uint8")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": warm, "ms_per_step": 1e3 * sum(walls) / len(walls),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a),
        "frame_pairs_per_step": workers,
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(a):
    """Identical for both arms (how many frames a step holds is reported beside it, not inside)."""
    return {"workload": "configs[3]: synthetic 4K VR180 stereo video, IDT frame-parallel",
            "frame": [a.height, a.width, 3], "bins": BINS, "n_iter": N_ITER,
            "distribution": "uniform-noise" if a.stress else "smooth-field+noise",
            "frames": "uint8 video frames decoded as k/255 float32 (utils/data.py:106); float32 frames for the device-resident value",
            "parallelism": f"frame-parallel x{a.gpus}, no collective",
            "l2": "working set per frame pair (200 MB float32 in, 199 MB state) exceeds the 126 MB L2"}


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
    H, W, F, P = a.height, a.width, a.frames, max(1, a.passes)
    npix = H * W
    warmup = max(a.warmup, 3)

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

    ctx = dict(torch=torch, dist=dist, device=device, synth=synth, _cabi=_cabi, batch=batch, handle=handle, dev=dev,
               world=world, rank=rank, barrier=barrier, max_over_ranks=max_over_ranks, peak=measured_peaks()[0])

    if a.linear_only:
        line = linear_section(a, ctx)
        if rank == 0:
            emit(line)
        return

    # frames k = rank, rank + world, ...: rotations pre-drawn in frame order after seed 42
    np.random.seed(42)
    all_rot = np.stack([batch.draw_rotations(N_ITER) for _ in range(F * world)])
    my_rot = torch.from_numpy(all_rot[rank::world].copy()).to(dev)
    tgt, ref = synth.frame_pairs_cuda(F, H, W, 2000 + rank, dev, stress=a.stress)
    out = torch.empty((F, H, W, 3), dtype=torch.float64, device=dev)
    ws = device.idt_workspace(npix, F, BINS, N_ITER, dev)

    def one_pass():
        device.idt_transfer(tgt, ref, my_rot, BINS, N_ITER, out=out, workspace=ws, handle=handle)

    def step():
        for _ in range(P):
            one_pass()

    for _ in range(warmup):
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
    value = world * F * P * npix / 1e6 / (ms_step / 1e3)
    if a.kernels_only:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "Mpix/s", "ms_per_step": ms_step, "ms_per_pass": ms_step / P,
                  "gpu_launches": int(launches), "note": "kernels-only profiling run"})
        return

    # ---- per-kernel times: CUDA events after every launch of the fused driver, on two extra steps of the
    # very loop timed above (ct_profile_*); reported per pass over the F resident frame pairs
    handle.profile(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step()
    step()
    p1.record()
    prof = handle.profile_read(16384)
    handle.profile(False)
    n_pass = 2 * P
    profiled_ms_per_pass = p0.elapsed_time(p1) / n_pass
    kern_ms, kern_launches = {}, {}
    for name, ms in prof:
        kind = name.split("_")[0]
        kern_ms[kind] = kern_ms.get(kind, 0.0) + ms / n_pass
        kern_launches[kind] = kern_launches.get(kind, 0) + 1
    # algorithmic bytes per pixel and pass of each kernel kind (float32 frames, fp64 state)
    kind_bytes = {"hist": (12 + 12) + 3 * (24 + 12), "remap": (12 + 24) + 3 * (24 + 24), "ranges": 12 + 12, "seed": 0}
    dominant = max((k for k in kern_ms if k in ("hist", "remap", "ranges")), key=lambda k: kern_ms[k])
    peak, peak_kind = measured_peaks()
    launches_per_pass = kern_launches[dominant] / n_pass
    alg_bytes_launch = kind_bytes[dominant] * npix * F / launches_per_pass
    avg_launch_ms = kern_ms[dominant] / launches_per_pass
    ach = alg_bytes_launch / (avg_launch_ms / 1e3) / 1e9
    kname = {"hist": "hist_kernel", "remap": "remap_kernel", "ranges": "ranges_kernel"}[dominant]
    bpp, bpp_src = ncu_traffic_bytes_per_pixel(kname)
    step_gbps = IDT_BYTES_PER_PIXEL_F32 * npix * F * P / (ms_step / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": peak, "peak_source": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "avg_launch_ms": avg_launch_ms,
                "launches_per_pass": launches_per_pass,
                "traffic": None if bpp is None else bpp * npix * F,
                "traffic_source": None if bpp is None else bpp_src + ": ncu dram bytes per pixel of the steady-state launches x pixels per launch",
                "kernel_ms_per_pass": {k: round(v, 4) for k, v in kern_ms.items()},
                "profiled_ms_per_pass": round(profiled_ms_per_pass, 4),   # the same two steps end to end: kernels + what lies between two passes
                "kernel_frac": {k: round(kind_bytes[k] * npix * F / (v / 1e3) / 1e9 / peak, 4) for k, v in kern_ms.items() if kind_bytes.get(k)},
                "timing": "CUDA events after every launch of the fused driver (ct_profile_*), two steps right after the timed region",
                "step_algorithmic_GBps": step_gbps, "step_frac": step_gbps / peak}
    del ws

    # ---- end to end, headline: uint8 video frames through the host API, pinned buffers, copies inside the timed region
    Fe = max(1, a.e2e_frames)
    e2e_steps = max(3, min(a.steps, 10))
    rot_e2e = np.stack([all_rot[(rank + world * k) % len(all_rot)] for k in range(Fe)])
    h8t = batch.pinned_empty((Fe, H, W, 3), np.uint8)
    h8r = batch.pinned_empty((Fe, H, W, 3), np.uint8)
    h8o = batch.pinned_empty((Fe, H, W, 3), np.uint8)
    t8 = (tgt * 255).round().to(torch.uint8).cpu().numpy()      # the resident frames are exact k/255
    r8 = (ref * 255).round().to(torch.uint8).cpu().numpy()
    for k in range(Fe):
        h8t[k] = t8[k % F]
        h8r[k] = r8[k % F]
    handle.set_stream(0)
    batch.idt_frames_u8(h8t, h8r, BINS, N_ITER, rotations=rot_e2e, out=h8o, handle=handle)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.idt_frames_u8(h8t, h8r, BINS, N_ITER, rotations=rot_e2e, out=h8o, handle=handle)
    barrier()
    u8_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e = {"value": world * Fe * npix / 1e6 / u8_s, "unit": "Mpix/s", "h2d_bytes_per_step": int(2 * Fe * npix * 3),
           "d2h_bytes_per_step": int(Fe * npix * 3), "frame_pairs_per_step": Fe, "steps": e2e_steps, "ms_per_step": u8_s * 1e3,
           "api": "color_transfer_b200.batch.idt_frames_u8 -> ct_idt_transfer_host_u8 (pinned uint8 frames in and out; "
                  "decode / encode fused into the kernels that read / write them)",
           "algorithmic_bytes_per_pixel": IDT_BYTES_PER_PIXEL_U8}
    # the uint8 path must give exactly the bytes of "device-resident float32 frames -> float64 -> quantise"
    want8 = np.rint(np.clip(out[0].cpu().numpy(), 0, 1) * 255).astype(np.uint8)
    e2e["matches_float_path_bytes"] = float(np.mean(h8o[0] == want8))
    del t8, r8

    # ---- the same through the float API: pinned float32 frames in, the reference's float64 result out
    Ff = min(8, F)
    ht = batch.pinned_empty((Ff, H, W, 3), np.float32)
    hr = batch.pinned_empty((Ff, H, W, 3), np.float32)
    ho = batch.pinned_empty((Ff, H, W, 3), np.float64)
    ht[...] = tgt[:Ff].cpu().numpy()
    hr[...] = ref[:Ff].cpu().numpy()
    rot_f = all_rot[rank::world][:Ff]
    batch.idt_frames(ht, hr, BINS, N_ITER, rotations=rot_f, out=ho, handle=handle)
    f_steps = max(2, min(a.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(f_steps):
        batch.idt_frames(ht, hr, BINS, N_ITER, rotations=rot_f, out=ho, handle=handle)
    barrier()
    f_s = max_over_ranks(time.perf_counter() - t0) / f_steps
    e2e_float = {"value": world * Ff * npix / 1e6 / f_s, "unit": "Mpix/s", "h2d_bytes_per_step": int(2 * Ff * npix * 12),
                 "d2h_bytes_per_step": int(Ff * npix * 24), "frame_pairs_per_step": Ff, "ms_per_step": f_s * 1e3,
                 "api": "color_transfer_b200.batch.idt_frames -> ct_idt_transfer_host (float32 in, float64 out)"}
    same = bool(np.array_equal(ho[0], out[0].cpu().numpy()))
    host_dma = host_dma_ceiling(ctx, ht, ho)
    # the video path moves 6 B/px in and 3 B/px out: its host-to-device stream against what the box gives all ranks
    e2e["h2d_GBps_aggregate"] = world * e2e["h2d_bytes_per_step"] / u8_s / 1e9
    e2e["frac_of_host_h2d"] = (e2e["h2d_GBps_aggregate"] / host_dma["h2d_only_GBps_aggregate"]) if "h2d_only_GBps_aggregate" in host_dma else None
    del ht, hr, ho, h8t, h8r, h8o

    # ---- device-resident uint8 frames (what the video path runs between the copies)
    t8d, r8d = (tgt * 255).round().to(torch.uint8), (ref * 255).round().to(torch.uint8)
    o8d = torch.empty_like(t8d)
    for _ in range(2):
        device.idt_transfer(t8d, r8d, my_rot, BINS, N_ITER, out=o8d, handle=handle)
    barrier()
    e0.record()
    for _ in range(P):
        device.idt_transfer(t8d, r8d, my_rot, BINS, N_ITER, out=o8d, handle=handle)
    e1.record()
    barrier()
    ms8 = max_over_ranks(e0.elapsed_time(e1)) / P
    value_u8 = {"value": world * F * npix / 1e6 / (ms8 / 1e3), "unit": "Mpix/s", "ms_per_pass": ms8,
                "frac_of_hbm": IDT_BYTES_PER_PIXEL_U8 * npix * F / (ms8 / 1e3) / 1e9 / peak,
                "note": "device-resident uint8 frames in and out, 243 algorithmic B/px"}
    del t8d, r8d, o8d, tgt, ref, out
    torch.cuda.empty_cache()

    linear = rowshard = extras = None
    if not a.no_extras:
        linear = linear_section(a, ctx)
        if world > 1:
            rowshard = rowshard_section(a, ctx)
        if rank == 0:
            extras = single_pair_section(ctx)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        np.random.seed(42)
        pool = CpuPool(1)
        vals, wall = [], 0.0
        for s in range(2):                       # two full frame pairs: ~12 s of CPU work
            v, w_, _ = pool.step(H, W, 2000 + s, draw_rotations_np(1))
            vals.append(v)
            wall += w_
        cpu = {"value": sum(vals) / len(vals), "unit": "Mpix/s", "cores": 1, "kind": "port",
               "sample": f"two {W}x{H} uint8 frame pairs of the workload (of the {F * P} per step): decode, numpy oracle port of "
                         f"methods/iterative.py (bins={BINS}, n_iter={N_ITER}), encode; single process ({wall:.1f} s)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps,
                "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a),
                "frame_pairs_per_step": F * P * world, "resident_frame_pairs_per_rank": F, "passes_per_step": P,
                "timed_region_s": ms_total / 1e3,
                "clocks": clocks, "e2e": e2e, "e2e_float": e2e_float, "host_dma": host_dma, "value_u8": value_u8,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "host_api_matches_device_api": same, "linear": linear, "rowshard": rowshard, "extras": extras}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def host_dma_ceiling(ctx, pinned_a, pinned_b):
    """What the host <-> device DMA of this box gives all ranks TOGETHER: every rank copies pinned host
    memory to its GPU (h2d_only), back (d2h_only), and both at once on two streams (both_*), no kernels.
    An end-to-end number at N GPUs cannot exceed bytes-per-pixel over this."""
    torch, dev = ctx["torch"], ctx["dev"]
    try:
        ha = torch.from_numpy(pinned_a.reshape(-1).view("uint8"))
        hb = torch.from_numpy(pinned_b.reshape(-1).view("uint8"))
        n = min(ha.numel(), hb.numel(), 1 << 29)
        ha, hb = ha[:n], hb[:n]
        da = torch.empty(n, dtype=torch.uint8, device=dev)
        db = torch.empty(n, dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        reps = 4

        def run(h2d, d2h):
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s_in):
                        da.copy_(ha, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s_out):
                        hb.copy_(db, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()

        def timed(h2d, d2h):
            run(h2d, d2h)
            ctx["barrier"]()
            t0 = time.perf_counter()
            run(h2d, d2h)
            ctx["barrier"]()
            dt = ctx["max_over_ranks"](time.perf_counter() - t0)
            return ctx["world"] * reps * n / dt / 1e9
        return {"h2d_only_GBps_aggregate": timed(True, False), "d2h_only_GBps_aggregate": timed(False, True),
                "both_GBps_aggregate_each_way": timed(True, True), "bytes_per_copy": int(n), "ranks": ctx["world"],
                "note": "pinned copies on every rank at the same time, no kernels; wall clock, max over ranks"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)}


def linear_section(a, ctx):
    """configs[2]: `--linear-pairs` (1035) pairs of 960x540 float32 per rank, frame-parallel (no collective):
    Reinhard f32 -> f32, MKL and CCS f32 -> f64, smooth-field and i.i.d.-uniform frames, ONE device-resident
    call per method; CUDA events, max over ranks; value = all ranks' pixels / that time."""
    torch, device, synth, _cabi, handle, dev = ctx["torch"], ctx["device"], ctx["synth"], ctx["_cabi"], ctx["handle"], ctx["dev"]
    world, rank, peak = ctx["world"], ctx["rank"], ctx["peak"]
    B, H, W = a.linear_pairs, 540, 960
    res = {"config": {"workload": "configs[2]: batch of synthetic 960x540 stereopairs, linear transfer, frame-parallel",
                      "pairs_per_rank": B, "shape": [H, W, 3], "input_dtype": "float32", "parallelism": f"frame-parallel x{world}, no collective",
                      "l2": "batch working set (6.4 GB per eye) exceeds the L2"},
           "unit": "Mpix/s"}
    methods = (("reinhard_f32", _cabi.CT_REINHARD, 48, torch.float32), ("mkl_f32_to_f64", _cabi.CT_MKL_MK, 60, torch.float64),
               ("ccs_f32_to_f64", _cabi.CT_CCS, 60, torch.float64))
    for dist_name, stress in (("smooth-field+noise", False), ("uniform-noise", True)):
        tgt, ref = synth.frame_pairs_cuda(B, H, W, 1000 + rank, dev, stress=stress)
        for name, method, bpp, odt in methods:
            if stress and method == _cabi.CT_CCS:
                continue
            dst = torch.empty((B, H, W, 3), dtype=odt, device=dev)
            for _ in range(2):
                device.linear_transfer(method, tgt, ref, out=dst, handle=handle)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx["barrier"]()
            reps = 5
            e0.record()
            for _ in range(reps):
                device.linear_transfer(method, tgt, ref, out=dst, handle=handle)
            e1.record()
            ctx["barrier"]()
            ms = ctx["max_over_ranks"](e0.elapsed_time(e1)) / reps
            gbps = bpp * B * H * W / (ms / 1e3) / 1e9           # per GPU
            res[f"{name}/{dist_name}"] = {"value": world * B * H * W / 1e6 / (ms / 1e3), "ms_per_batch": ms,
                                          "algorithmic_bytes_per_pixel": bpp, "per_gpu_GBps": gbps, "frac_of_hbm": gbps / peak}
            del dst
        del tgt, ref
        torch.cuda.empty_cache()
    res["value"] = res["reinhard_f32/smooth-field+noise"]["value"]
    res["metric"] = "stereopair Mpix/s (Reinhard, 1035 x 960x540 float32 pairs per rank, frame-parallel)"
    return res


def rowshard_section(a, ctx):
    """configs[4]: one side x side float32 pair, contiguous row blocks per rank (strong scaling).
    IDT: 1 + n_iter MIN all-reduces of 6 int64 keys and n_iter SUM all-reduces of 6*bins int64 counts;
    MKL / Reinhard: one all-gather of 2x10 raw moments.  Device-resident shards, CUDA events, max over
    ranks.  Before the timing, the same drivers run on a 4096x4096 pair of the same generator and every
    rank compares its rows with the UNSHARDED result it computes itself: bit-identical histogram counts and
    output for IDT, <= 1e-12 for MKL, <= 2e-6 (float32) for Reinhard."""
    import numpy as np
    from color_transfer_b200 import sharded
    torch, dist, device, synth, _cabi, handle, dev = (ctx[k] for k in ("torch", "dist", "device", "synth", "_cabi", "handle", "dev"))
    world, rank, peak = ctx["world"], ctx["rank"], ctx["peak"]
    comm = sharded.Comm()
    np.random.seed(42)
    rot = sharded.predraw_rotations(1, N_ITER)[0]
    drot = torch.from_numpy(rot[None]).to(dev)

    # ---- parity: 4096 x 4096, every rank holds the whole pair and its own row block
    side_p = 4096
    tp, rp = synth.frame_pairs_cuda(1, side_p, side_p, 3000, dev)
    tp, rp = tp[0], rp[0]
    p0, p1 = sharded.row_partition(side_p, world, rank)
    full_counts = []
    st = device.IdtStages(tp, rp, drot, BINS, N_ITER, handle=handle)
    full = st.run(between=lambda n, x: full_counts.append(x.clone()) if n == "counts" else None, fuse_lut=False)[0]
    part_counts = []

    def between(name, tensor):
        if name == "keys":
            comm.min_(tensor)
        else:
            comm.sum_(tensor)
            part_counts.append(tensor.clone())

    backend = sharded.CudaIdtBackend(tp[p0:p1].contiguous(), rp[p0:p1].contiguous(), rot, BINS, N_ITER, handle)
    part = backend.run(between)
    backend.finish()
    counts_ok = len(part_counts) == len(full_counts) and all(torch.equal(x, y) for x, y in zip(part_counts, full_counts))
    idt_ok = bool(torch.equal(part, full[p0:p1]))
    mkl = sharded.linear_transfer_sharded(_cabi.CT_MKL_MK, tp[p0:p1].contiguous(), rp[p0:p1].contiguous(), comm=comm, handle=handle)
    mkl_err = float((mkl - device.linear_transfer(_cabi.CT_MKL_MK, tp, rp, handle=handle)[p0:p1]).abs().max())
    rei = sharded.linear_transfer_sharded(_cabi.CT_REINHARD, tp[p0:p1].contiguous(), rp[p0:p1].contiguous(), comm=comm, handle=handle)
    rei_err = float((rei - device.linear_transfer(_cabi.CT_REINHARD, tp, rp, handle=handle)[p0:p1]).abs().max())
    ok_local = counts_ok and idt_ok and mkl_err <= 1e-12 and rei_err <= 2e-6
    flag = torch.tensor([1 if ok_local else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    errs = torch.tensor([mkl_err, rei_err], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    parity = {"parity_ok": bool(flag.item() == 1), "pair": [side_p, side_p, 3], "idt_counts_bit_identical": counts_ok,
              "idt_output_bit_identical": idt_ok, "mkl_max_abs": float(errs[0]), "reinhard_f32_max_abs": float(errs[1]),
              "against": "the unsharded single-GPU result, computed by every rank"}
    del tp, rp, full, part, mkl, rei, st, backend
    torch.cuda.empty_cache()

    # ---- timing: side x side, every rank generates only its rows
    side = a.rowshard_side
    r0, r1 = sharded.row_partition(side, world, rank)
    rows = r1 - r0
    tgt, ref = synth.frame_pairs_cuda(1, rows, side, 3000 + r0, dev)
    tgt, ref = tgt[0], ref[0]
    npix = side * side

    def timed(fn, reps):
        for _ in range(2):
            fn()
        ctx["barrier"]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        ctx["barrier"]()
        return ctx["max_over_ranks"](e0.elapsed_time(e1)) / reps

    backend = sharded.CudaIdtBackend(tgt, ref, rot, BINS, N_ITER, handle)
    calls0 = comm.calls
    backend.run(lambda name, tensor: comm.min_(tensor) if name == "keys" else comm.sum_(tensor))
    collectives_idt = comm.calls - calls0
    ms = timed(lambda: backend.run(lambda name, tensor: comm.min_(tensor) if name == "keys" else comm.sum_(tensor)), 5)
    res = {"config": {"workload": f"configs[4]: single {side}x{side} float32 stereopair row-sharded across {world} GPUs with NCCL "
                                  "all-reduce of range keys / histogram counts / moments",
                      "rows_per_rank": rows, "bins": BINS, "n_iter": N_ITER},
           "scaling": "strong", "unit": "Mpix/s", "parity": parity, "parity_ok": parity["parity_ok"],
           "idt": {"ms_per_pair": ms, "value": npix / 1e6 / (ms / 1e3), "collectives": collectives_idt,
                   "frac_of_hbm_aggregate": IDT_BYTES_PER_PIXEL_F32 * npix / (ms / 1e3) / 1e9 / (peak * world)}}
    del backend
    for name, code, bpp in (("mkl", _cabi.CT_MKL_MK, 60), ("reinhard", _cabi.CT_REINHARD, 48)):
        ms = timed(lambda: sharded.linear_transfer_sharded(code, tgt, ref, comm=comm, handle=handle), 5)
        res[name] = {"ms_per_pair": ms, "value": npix / 1e6 / (ms / 1e3), "collectives": 1,
                     "frac_of_hbm_aggregate": bpp * npix / (ms / 1e3) / 1e9 / (peak * world)}
    # what the 8 collectives of one IDT cost by themselves: the two exchanges of an iteration on empty streams
    keys = torch.zeros(6, dtype=torch.int64, device=dev)
    counts = torch.zeros(2 * 3 * BINS, dtype=torch.int64, device=dev)
    us_pair = timed(lambda: (comm.min_(keys), comm.sum_(counts)), 50) * 1e3
    res["idt"]["collective_us"] = {"min_keys_plus_sum_counts": us_pair, "per_pair_total": us_pair * collectives_idt / 2,
                                   "note": "NCCL all-reduce MIN of 6 int64 + SUM of 6*bins int64, back to back on an idle stream"}
    res["ms_per_pair"] = res["idt"]["ms_per_pair"]
    res["frac_of_hbm_aggregate"] = res["idt"]["frac_of_hbm_aggregate"]
    res["collectives"] = collectives_idt
    del tgt, ref
    torch.cuda.empty_cache()
    return res


def single_pair_section(ctx):
    """configs[0] / configs[1] shape: one 1080x860 float64 pair (L2-resident, launch-latency regime) on the
    device, and the reference's own stereo pair through the reference-facing numpy functions."""
    import numpy as np
    torch, device, synth, _cabi, batch, handle, dev = (ctx[k] for k in ("torch", "device", "synth", "_cabi", "batch", "handle", "dev"))
    out = {}
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
    out["distortion_grid_4k"] = distortion_section(ctx)
    return out


def distortion_section(ctx):
    """the 31 test-set distortions (ref: utils/data.py:12-22) of one 3840x2160 uint8 frame in one
    pass, against the CPU oracle port on a 1/16-area crop (one thread)."""
    import numpy as np
    torch, dev = ctx["torch"], ctx["dev"]
    from color_transfer_b200 import data
    from oracle import distort_numpy
    H, W = 2160, 3840
    img = torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev)
    fns = data.setup_grid_distortions()
    for _ in range(3):
        got = data.distort_grid(img, fns)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        got = data.distort_grid(img, fns)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    crop = img[:, :H // 4, :W // 4].contiguous()
    got_crop = data.distort_grid(crop, fns).cpu().numpy()
    t0 = time.perf_counter()
    want = distort_numpy.distort_grid(crop.cpu().numpy())
    cpu_ms = (time.perf_counter() - t0) * 1e3 * 16
    diff = np.abs(got_crop.astype(np.int16) - want.astype(np.int16))
    gbps = H * W * 3 * (1 + len(fns)) / ms / 1e6
    return {"distortions": len(fns), "ms_per_frame": ms, "Mpix_copies/s": H * W * len(fns) / ms / 1e3,
            "algorithmic_bytes_per_pixel": 3 * (1 + len(fns)), "GBps": gbps, "frac_of_hbm": gbps / measured_peaks()[0],
            "bound": "instruction issue (hue: 6 IEEE divisions per pixel; table look-ups in shared memory), not HBM",
            "cpu_oracle_ms_per_frame": cpu_ms, "cpu_sample": "1/16-area crop, numpy port, one thread, scaled by 16",
            "max_level_diff_vs_oracle": int(diff.max()), "identical_fraction": float((diff == 0).mean())}


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

"""Multi-rank body of the row-sharded / frame-parallel GPU checks.  Prints DIST_GPU_CHECK_OK on rank 0.

Two ways to launch it:

* one GPU per rank over NCCL (the product configuration):
      python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/dist_gpu_check.py
* every rank on cuda:0 over gloo (CT_DIST_SAME_DEVICE=1): the same drivers, kernels and collective
  schedule on a one-GPU box - NCCL refuses two ranks on one device, gloo all-reduces CUDA tensors
  through the host.  This is what the driver's one-GPU `pytest -m gpu` run exercises.

Checks:
  1. 257x193 pair, ragged reference: row-sharded IDT == single GPU (bit-identical counts, output <1e-12).
  2. config-5 protocol at a size the CPU oracle finishes in seconds: a SIDE x SIDE float32 pair of the
     config-5 generator (seed 3000), rows sharded, against (a) the single-GPU run - bit-identical
     counts AND bit-identical output - and (b) the CPU oracle on rank 0 - bit-exact counts, <=1e-9.
  3. row-sharded linear transfers vs single GPU.
  4. frame-parallel IDT bit-identical to per-frame calls.
A failing rank writes its traceback to gpurun_out/dist_rank{r}.txt before exiting non-zero.
"""

import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import color_transfer_b200  # noqa: E402,F401
from color_transfer_b200 import _cabi, device, sharded, synth  # noqa: E402
from conftest import synthetic_pair  # noqa: E402

SIDE = int(os.environ.get("CT_DIST_SIDE", "2048"))


def sharded_idt_with_counts(comm, dt, dr, rot, dev):
    backend = sharded.CudaIdtBackend(dt, dr, rot, 255, 4)
    log = []

    def between(name, tensor):
        if name == "keys":
            comm.min_(tensor)
        else:
            comm.sum_(tensor)
            log.append(tensor.clone())

    part = backend.run(between)
    backend.finish()
    return part, log


def single_idt_with_counts(dt, dr, rot, dev):
    log = []
    st = device.IdtStages(dt, dr, torch.from_numpy(rot[None]).to(dev))
    full = st.run(between=lambda n, x: log.append(x.clone()) if n == "counts" else None, fuse_lut=False)[0]
    st.raise_for_status()
    return full, log


def body(rank, world, dev):
    comm = sharded.Comm()
    rot = sharded.predraw_rotations(1, 4, seed=42)[0]

    # ---- 1. small ragged pair: sharded vs single GPU (every rank computes the full problem as the check)
    t, r = synthetic_pair(257, 193, 77, np.float32, ref_shape=(201, 160))
    dt, dr = torch.from_numpy(t).to(dev), torch.from_numpy(r).to(dev)
    a, b = sharded.row_partition(t.shape[0], world, rank)
    ra, rb = sharded.row_partition(r.shape[0], world, rank)
    full, single_counts = single_idt_with_counts(dt, dr, rot, dev)
    part, counts_log = sharded_idt_with_counts(comm, dt[a:b].contiguous(), dr[ra:rb].contiguous(), rot, dev)
    err = float((part - full[a:b]).abs().max())
    assert err < 1e-12, f"row-sharded IDT differs from single GPU: {err}"
    for i, (c_sharded, c_single) in enumerate(zip(counts_log, single_counts)):
        assert torch.equal(c_sharded, c_single), f"sharded histogram counts are not bit-identical (iteration {i})"

    # ---- 2. config-5 protocol, SIDE x SIDE
    bt, br = synth.frame_pair(SIDE, SIDE, 3000, np.float32)
    ba, bb = sharded.row_partition(SIDE, world, rank)
    dbt, dbr = torch.from_numpy(bt).to(dev), torch.from_numpy(br).to(dev)
    big_full, big_single_counts = single_idt_with_counts(dbt, dbr, rot, dev)
    big_part, big_counts = sharded_idt_with_counts(comm, dbt[ba:bb].contiguous(), dbr[ba:bb].contiguous(), rot, dev)
    for i, (c_sharded, c_single) in enumerate(zip(big_counts, big_single_counts)):
        nd = int((c_sharded != c_single).sum())
        assert nd == 0, f"{SIDE}x{SIDE}: sharded counts differ from single GPU at iteration {i} in {nd} bins"
    assert torch.equal(big_part, big_full[ba:bb]), \
        f"{SIDE}x{SIDE}: sharded output is not bit-identical to single GPU: {float((big_part - big_full[ba:bb]).abs().max())}"
    big_err = -1.0
    if rank == 0:
        from oracle import reference_numpy as oracle
        want, traces = oracle.idt_instrumented(bt, br, rotations=rot, keep_arrays=False)
        for i, c in enumerate(big_counts):
            c = c.cpu().numpy().reshape(2, 3, 255)
            dt_ = int(np.abs(c[0] - traces[i]["counts_t"]).sum())
            dr_ = int(np.abs(c[1] - traces[i]["counts_r"]).sum())
            assert dt_ == 0 and dr_ == 0, f"row-sharded counts differ from the oracle at iteration {i}: L1 {dt_} / {dr_}"
        big_err = float(np.max(np.abs(big_part.cpu().numpy() - want[ba:bb])))
        assert big_err < 1e-9, f"row-sharded {SIDE}x{SIDE} IDT differs from the oracle: {big_err}"
    del dbt, dbr, big_full, big_part

    # ---- 3. linear: sharded vs single GPU
    for code in (_cabi.CT_MKL_MK, _cabi.CT_REINHARD):
        want = device.linear_transfer(code, dt, dr)[a:b]
        got = sharded.linear_transfer_sharded(code, dt[a:b].contiguous(), dr[ra:rb].contiguous(), comm=comm)
        rel = float((got.double() - want.double()).abs().max())
        assert rel < 1e-6 if code == _cabi.CT_REINHARD else rel < 1e-12, f"sharded linear differs: {rel}"

    # ---- 4. frame-parallel: frame k on rank k mod world, no collective on the data path
    frames = 5
    rots = sharded.predraw_rotations(frames, 4, seed=7)
    mine = sharded.frame_partition(frames, world, rank)
    if mine:
        pairs = [synthetic_pair(64, 96, 500 + k, np.float32) for k in mine]
        ft = torch.from_numpy(np.stack([p[0] for p in pairs])).to(dev)
        fr = torch.from_numpy(np.stack([p[1] for p in pairs])).to(dev)
        out = device.idt_transfer(ft, fr, torch.from_numpy(rots[mine]).to(dev))
        for i, k in enumerate(mine):
            single = device.idt_transfer(ft[i], fr[i], torch.from_numpy(rots[k][None]).to(dev))
            assert torch.equal(out[i], single)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_CHECK_OK", f"world={world}", f"idt_err={err:.2e}", f"oracle_err_{SIDE}={big_err:.2e}",
              f"collectives={comm.calls}", flush=True)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    same_device = os.environ.get("CT_DIST_SAME_DEVICE") == "1"
    local = 0 if same_device else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    try:
        body(rank, world, dev)
    except BaseException:  # noqa: BLE001
        tb = traceback.format_exc()
        sys.stderr.write(f"[rank {rank}] {tb}\n")
        out_dir = os.path.join(ROOT, "gpurun_out")
        try:
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, f"dist_rank{rank}.txt"), "w") as f:
                f.write(tb)
        except OSError:
            pass
        os._exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

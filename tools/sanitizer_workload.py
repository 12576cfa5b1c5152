import sys; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import synthetic_pair
import methods.linear as lin, methods.iterative as it
from color_transfer_b200 import batch
for dtype in (np.float32, np.float64):
    t, r = synthetic_pair(67, 93, 3, dtype, (50, 81))
    lin.color_transfer_between_images(t, r); lin.monge_kantorovitch_color_transfer(t, r, "cholesky"); lin.color_transfer_in_correlated_color_space(t, r)
    np.random.seed(1); it.iterative_distribution_transfer(t, r)
    it.iterative_distribution_transfer(t, r, bins=1024, n_iter=2)
    tc = np.ascontiguousarray(t.transpose(2,0,1)).transpose(1,2,0); it.iterative_distribution_transfer(tc, r, n_iter=2)
t, r = synthetic_pair(70, 130, 5, np.float32)   # > one f32 tile (1024 px): exercises the TMA pipeline
np.random.seed(2); it.iterative_distribution_transfer(t, r)
t8 = np.rint(np.stack([t, t]) * 255).astype(np.uint8); r8 = np.rint(np.stack([r, r]) * 255).astype(np.uint8)
batch.idt_frames_u8(t8, r8); batch.linear_transfer_frames_u8("reinhard", t8, r8)
t, r = synthetic_pair(97, 131, 7, np.float64)
np.random.seed(3); it.automated_color_grading(t, r)
# round-1 second session: fp32 Lab chain incl. the patched toes, chunked two-stream batches, metrics
import torch
from color_transfer_b200 import _cabi, device, metrics
rng = np.random.default_rng(9)
dark = (rng.integers(0, 14, (35, 3, 23, 31)) / 255.0).astype(np.float32)
dark[::3, :, ::4, ::5] = rng.random(dark[::3, :, ::4, ::5].shape, dtype=np.float32)
tb = torch.from_numpy(np.ascontiguousarray(dark.transpose(0, 2, 3, 1))).cuda()
device.linear_transfer(_cabi.CT_REINHARD, tb, tb.flip(0)); device.linear_transfer(_cabi.CT_MKL_MK, tb, tb.flip(0))
x = torch.rand(2, 3, 75, 101, device="cuda"); y = (x * 0.8 + 0.1).clamp(0, 1)
metrics.icid(x, y); metrics.icid(x, y, downsampling=False, omit_maps67=True); metrics.psnr(x, y); metrics.ssim(x, y)
big = torch.rand(1, 3, 530, 300, device="cuda"); metrics.icid(big, big.flip(3)); metrics.ssim(big, big.flip(3))
torch.cuda.synchronize()
print("sanitizer workload done")
# round 2: uint8 frames decoded / encoded inside the kernels (interleaved and planar, float32 / uint8 / clamped results),
# the seeded fp32 screens of K4 / K5 (6 rotations = two K4 chunks), the split-rotation K4 variant, scalar tails
t8d = torch.from_numpy(t8).cuda(); r8d = torch.from_numpy(r8).cuda()
rot6 = torch.from_numpy(np.stack([batch.draw_rotations(6) for _ in range(2)])).cuda()
device.idt_transfer(t8d, r8d, rot6, 255, 6)
device.idt_transfer(t8d, r8d, rot6[:, :4].contiguous(), 255, 4, out_dtype=torch.float32, clamp=True)
t8p = t8d.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
device.idt_transfer(t8p, r8d, rot6[:, :2].contiguous(), 64, 2, as_float32=False)
for code in (_cabi.CT_REINHARD, _cabi.CT_MKL_MK):
    device.linear_transfer(code, t8p, r8d); device.linear_transfer(code, t8d, r8d, out_dtype=torch.float32, clamp=True)
odd = torch.rand(3, 33, 37, 3, device="cuda", dtype=torch.float64)
device.idt_transfer(odd, odd.flip(0), rot6[:1, :4].expand(3, 4, 3, 3).contiguous(), 255, 4)
torch.cuda.synchronize()
print("round-2 sanitizer workload done")
# round 2, second session: 256-bit stores (32-byte aligned destinations: float32 pairs -> float64 state / results, the
# float64 interleaved result of the last iteration), the last-releaser tile refill, the distortion generator
al32 = torch.rand(2, 64, 96, 3, device="cuda", dtype=torch.float32)
device.idt_transfer(al32, al32.flip(0), rot6[:, :4].contiguous(), 255, 4)
device.linear_transfer(_cabi.CT_MKL_MK, al32, al32.flip(0)); device.linear_transfer(_cabi.CT_CCS, al32, al32.flip(0))
al64 = al32.double()
device.idt_transfer(al64, al64.flip(0), rot6[:, :4].contiguous(), 255, 4); device.linear_transfer(_cabi.CT_MKL_MK, al64, al64.flip(0))
from color_transfer_b200 import data
fns = data.setup_grid_distortions()
img8 = torch.randint(0, 256, (3, 64, 96), dtype=torch.uint8, device="cuda")
data.distort_grid(img8, fns); data.distort_grid(img8[:, :33, :37].contiguous(), fns)
data.distort_grid(torch.stack([img8, img8.flip(1)]), fns[:7])
torch.cuda.synchronize()
print("round-2 (second session) sanitizer workload done")

import sys, numpy as np, torch
sys.path.insert(0, '.')
import color_transfer_b200
from color_transfer_b200 import device, synth, batch, _cabi
dev = torch.device('cuda:0')
F = 4
t, r = synth.frame_pairs_cuda(F, 2160, 3840, 100, dev, dtype=torch.float32)
np.random.seed(42)
rot = torch.from_numpy(np.stack([batch.draw_rotations(4) for _ in range(F)])).to(dev)
h = _cabi.default_handle(0)
out = torch.empty((F, 2160, 3840, 3), dtype=torch.float64, device=dev)
for _ in range(3): device.idt_transfer(t, r, rot, 255, 4, out=out, handle=h)
torch.cuda.synchronize()
h.profile(True)
for _ in range(5): device.idt_transfer(t, r, rot, 255, 4, out=out, handle=h)
torch.cuda.synchronize()
res = h.profile_read()
h.profile(False)
per = len(res) // 5
acc = np.zeros(per)
for k in range(5):
    for i in range(per): acc[i] += res[k * per + i][1]
for i in range(per): print(f"{res[i][0]:>18s} {acc[i]/5*1e3:8.1f} us")

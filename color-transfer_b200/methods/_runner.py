"""``Runner(func_spec)`` with the reference's forward contract (ref: methods/__init__.py:10-27):
a batch dict of CHW tensors in, a stacked CHW float32 tensor out.  It is a LightningModule when
pytorch_lightning is importable (so LightningCLI can drive it as in the reference) and a plain
torch.nn.Module otherwise.  ``test_step`` computes the two quality metrics that have a device
implementation here (PSNR, SSIM and iCID, SURVEY.md section 8f-3, ``color_transfer_b200.metrics``);
the reference's FSIM (piq) is not provided.

Fast path: when the batch tensors already live on a CUDA device and the
resolved function is one of this package's transfers, the whole batch is handed to the kernels
as ONE device-resident call - no ``.cpu().numpy()`` round trip, no per-pair Python loop.  The CHW
tensors are viewed as HWC without a copy (the kernels read planar images natively), uint8 tensors
(``torchvision.io.read_image`` output, ref: utils/data.py:99-106) are decoded as k/255 in float32
INSIDE the kernels that read them, and the ``.float()`` cast (and ``test_step``'s ``.clamp(0, 1)``)
is done by the kernel that writes the result."""

import torch

from . import resolve

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    _Base = torch.nn.Module


class Runner(_Base):
    def __init__(self, func_spec):
        super().__init__()
        self.func = resolve(func_spec)
        self._clamp_fused = False   # test_step: let the kernel that writes the result clamp it

    def forward(self, batch):
        target, reference = batch["target"], batch["reference"]
        device_impl = getattr(self.func, "device_impl", None)
        if (device_impl is not None and torch.is_tensor(target) and torch.is_tensor(reference)
                and target.is_cuda and reference.is_cuda and target.dim() == 4 and reference.dim() == 4):
            out = device_impl(target.permute(0, 2, 3, 1), reference.permute(0, 2, 3, 1),   # [B,H,W,3] views
                              out_dtype=torch.float32, clamp=self._clamp_fused)
            return out.permute(0, 3, 1, 2)
        outputs = []
        for t, r in zip(target, reference):
            # same marshalling as the reference: HWC *views* of CHW memory, float32
            t = t.permute(1, 2, 0).detach().cpu().numpy()
            r = r.permute(1, 2, 0).detach().cpu().numpy()
            output = torch.from_numpy(self.func(t, r)).float().permute(2, 0, 1)
            outputs.append(output)
        return torch.stack(outputs)

    def test_step(self, batch, batch_idx, dataloader_idx=0):
        """ref: methods/__init__.py:29-40 - clamp the result and score it against batch["gt"].
        PSNR, SSIM and iCID are computed on the device; returns (and, under Lightning, logs) them."""
        from ..metrics import icid, psnr, ssim
        self._clamp_fused = True
        try:
            result = self(batch).clamp(0, 1)   # a no-op copy after the fused clamp; kept for the host path
        finally:
            self._clamp_fused = False
        gt = batch["gt"]
        if not gt.is_cuda:
            gt = gt.cuda()
        result = result.to(gt.device)
        values = {"Test PSNR": psnr(result, gt), "Test SSIM": ssim(result, gt), "Test iCID": icid(result, gt)}
        if hasattr(self, "log"):
            for name, value in values.items():
                self.log(name, value, prog_bar=name == "Test PSNR")
        return values

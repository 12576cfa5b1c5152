"""``Runner(func_spec)`` with the reference's forward contract (ref: methods/__init__.py:10-27):
a batch dict of CHW tensors in, a stacked CHW float32 tensor out.  It is a LightningModule when
pytorch_lightning is importable (so LightningCLI can drive it as in the reference) and a plain
torch.nn.Module otherwise.  The quality metrics of ``test_step`` (piq, iCID) are outside the hot
path and are not provided here."""

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

    def forward(self, batch):
        outputs = []
        for target, reference in zip(batch["target"], batch["reference"]):
            # same marshalling as the reference: HWC *views* of CHW memory, float32
            target = target.permute(1, 2, 0).detach().cpu().numpy()
            reference = reference.permute(1, 2, 0).detach().cpu().numpy()
            output = torch.from_numpy(self.func(target, reference)).float().permute(2, 0, 1)
            outputs.append(output)
        return torch.stack(outputs)

"""Drop-in for ref: utils/icid.py — ``icid(img1, img2, intent, omit_maps67, downsampling)`` on CUDA
tensors, computed by libct_b200.so (color-transfer_b200/csrc/ct_metrics.cu)."""
import color_transfer_b200  # noqa: F401  (registers the hyphenated directory)
from color_transfer_b200.metrics import icid  # noqa: F401

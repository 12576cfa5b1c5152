"""Drop-in for the test-set half of ref: utils/data.py - ``setup_grid_distortions`` and the two test
datasets on CUDA tensors, computed by libct_b200.so (color-transfer_b200/csrc/ct_distort.cu).  The
training dataset and the Lightning ``DataModule`` belong to the neural methods and are out of scope."""
import color_transfer_b200  # noqa: F401  (registers the hyphenated directory)
from color_transfer_b200.data import (ArtificialTestDataset, RealWorldTestDataset, distort_grid,  # noqa: F401
                                      read_image, setup_grid_distortions)

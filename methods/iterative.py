"""ref: methods/iterative.py - ``iterative_distribution_transfer``, served by color-transfer_b200."""
from color_transfer_b200.methods.iterative import iterative_distribution_transfer  # noqa: F401

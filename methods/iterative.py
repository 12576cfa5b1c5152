"""ref: methods/iterative.py - ``iterative_distribution_transfer`` and ``automated_color_grading``,
served by color-transfer_b200."""
from color_transfer_b200.methods.iterative import (  # noqa: F401
    automated_color_grading,
    iterative_distribution_transfer,
)

"""ref: methods/linear.py - same three function names, served by color-transfer_b200."""
from color_transfer_b200.methods.linear import (  # noqa: F401
    color_transfer_between_images,
    color_transfer_in_correlated_color_space,
    monge_kantorovitch_color_transfer,
)

"""color_transfer_b200 - B200 (sm_100a) implementation of the global statistical colour-transfer
path of egorchistov/color-transfer.

Import name: ``color_transfer_b200`` (the directory is ``color-transfer_b200/``; the repo-root
module ``color_transfer_b200.py`` registers it, or put this directory's parent on sys.path via
that shim).  The drop-in mirror of the reference's ``methods`` package is
``color_transfer_b200.methods`` (also re-exported by the repo-root ``methods/`` package).
"""

from . import _cabi  # noqa: F401
from ._cabi import CtError, Handle, default_handle  # noqa: F401

__all__ = ["CtError", "Handle", "default_handle"]

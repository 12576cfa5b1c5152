"""Repo-root ``methods`` package: lets the reference's own entry points
(``--model.func_spec methods.linear.color_transfer_between_images``, the demo notebook's
``from methods.linear import ...``) resolve to the B200 implementation when this repository is
first on ``sys.path``.  Everything lives in ``color-transfer_b200/methods``."""
import color_transfer_b200  # noqa: F401  (registers the hyphenated directory)
from color_transfer_b200.methods import iterative, linear, resolve  # noqa: F401


def __getattr__(name):
    if name == "Runner":
        from color_transfer_b200.methods import Runner
        return Runner
    raise AttributeError(name)

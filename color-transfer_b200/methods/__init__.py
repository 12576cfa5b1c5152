"""Drop-in mirror of the reference's ``methods`` package for the statistical transfers.

ref: methods/__init__.py resolves ``func_spec`` (a dotted path such as
``methods.linear.color_transfer_between_images``, configs/others.yaml:5) with importlib and calls
``func(target, reference)`` on ``[H,W,3]`` numpy arrays.  ``Runner`` keeps that contract;
it is built lazily so that importing ``methods.linear`` never needs pytorch_lightning.
"""

import importlib

from . import iterative, linear  # noqa: F401

__all__ = ["linear", "iterative", "resolve", "Runner"]


def resolve(func_spec):
    """The callable named by a dotted path (ref: methods/__init__.py:14-16)."""
    parts = func_spec.split(".")
    module, func = ".".join(parts[:-1]), parts[-1]
    return importlib.import_module(module).__getattribute__(func)


def __getattr__(name):
    if name == "Runner":
        from ._runner import Runner
        return Runner
    raise AttributeError(name)

"""Repo-root ``utils`` package: only the pieces of the reference's ``utils`` that sit next to the
hot path and have a device implementation here (``utils.icid``).  Lets ``from utils.icid import
icid`` (ref: methods/__init__.py:7) resolve to the B200 implementation when this repository is
first on ``sys.path``."""

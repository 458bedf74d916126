from . import lib, ops  # noqa: F401

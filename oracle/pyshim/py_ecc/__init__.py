from . import bn128  # noqa: F401

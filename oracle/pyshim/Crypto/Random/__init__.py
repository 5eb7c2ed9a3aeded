import os
from . import random  # noqa: F401


def get_random_bytes(n):
    return os.urandom(n)

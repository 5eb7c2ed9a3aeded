import random as _r

_sys = _r.SystemRandom()


def randrange(*a):
    return _sys.randrange(*a)


def randint(a, b):
    return _sys.randint(a, b)

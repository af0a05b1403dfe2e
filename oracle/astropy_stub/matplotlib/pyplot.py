from . import _any


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _any

"""Permissive matplotlib stand-in: the reference imports pyplot at module level and its validation
helpers draw comparison plots inside tests; every attribute access / call here is a no-op."""


class _Anything:
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, key):
        return self

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter((self, self))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_any = _Anything()
rcParams = {}


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _any

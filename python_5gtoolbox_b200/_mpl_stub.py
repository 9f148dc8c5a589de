"""A do-nothing stand-in for matplotlib, for boxes without it: the reference's LDPC scripts import
``matplotlib.pyplot`` at module level (scripts/internal/sim_ldpc_internal.py:3) and draw a BLER figure after the
simulation; with the stub every plotting call is accepted and ignored, the pickle output is unaffected."""
import sys
import types


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter(())


def install():
    """Put the stub into sys.modules unless a real matplotlib is importable.  Returns True when the stub is in place."""
    try:
        import matplotlib.pyplot  # noqa: F401
        return False
    except ImportError:
        pass
    top = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    def anything(name):
        if name.startswith("__"):   # module protocol probes (__file__, __spec__, __wrapped__ ...) must keep failing
            raise AttributeError(name)
        return _Anything()
    plt.__getattr__ = anything
    top.__getattr__ = anything
    top.pyplot = plt
    top.__path__ = []
    sys.modules["matplotlib"] = top
    sys.modules["matplotlib.pyplot"] = plt
    return True

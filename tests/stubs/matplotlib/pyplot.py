"""pyplot test double: records plotted series, draws nothing (see the package docstring)."""
import json
import os


def _record(kind, args, kwargs):
    path = os.environ.get("MPL_STUB_LOG")
    if not path:
        return
    series = [[float(v) for v in a] for a in args if hasattr(a, "__len__") and not isinstance(a, str)]
    with open(path, "a") as f:
        f.write(json.dumps({"kind": kind, "series": series, "label": str(kwargs.get("label"))}) + "\n")


class _Anything:
    """An object every attribute of which is a callable returning another such object (axis.set_ticks, ax.legend, ...)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *args, **kwargs):
        return _Anything()


class _Axes(_Anything):
    def plot(self, *args, **kwargs):
        _record("plot", args, kwargs)
        return [_Anything()]


class _Figure(_Anything):
    def savefig(self, path, *args, **kwargs):
        _record("savefig", (), {"label": path})


class _PropCycle:
    def by_key(self):
        return {"color": ["C0", "C1", "C2"]}


rcParams = {"axes.prop_cycle": _PropCycle()}
_current = _Axes()


def subplots(nrows=1, ncols=1, **kwargs):
    n = nrows * ncols
    return _Figure(), (_Axes() if n == 1 else [_Axes() for _ in range(n)])


def plot(*args, **kwargs):
    return _current.plot(*args, **kwargs)


def savefig(path, *args, **kwargs):
    _record("savefig", (), {"label": path})


def __getattr__(name):        # xlabel, ylabel, legend, figure, close, ...
    return _Anything()

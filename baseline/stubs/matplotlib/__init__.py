"""Stand-in for matplotlib (absent from this image): lets `analysis.report` of the reference import; never called."""
import sys
import types


def use(*a, **k):
    return None


for _sub in ("pyplot", "ticker", "colors", "patches", "gridspec"):
    _m = types.ModuleType(__name__ + "." + _sub)
    sys.modules[__name__ + "." + _sub] = _m
    globals()[_sub] = _m

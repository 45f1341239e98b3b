import numpy as np

from . import units as _u


class SimArray(np.ndarray):
    def __new__(cls, data, units=None):
        obj = np.asarray(data, dtype=np.float64).view(cls)
        obj.units = _u.parse(units) if units is not None else _u.Unit(1.0, "1")
        obj.sim = None
        return obj

    def __array_finalize__(self, obj):
        self.units = getattr(obj, "units", _u.Unit(1.0, "1"))
        self.sim = getattr(obj, "sim", None)

    def in_units(self, new):
        new = _u.parse(new)
        out = SimArray(np.asarray(self) * self.units.ratio(new), new)
        out.sim = self.sim  # pynbody keeps the snapshot reference through unit conversions
        return out

"""Shear-function hooks — same classes, arguments and error behaviour as PSEv1/shear_function.py:10-114;
the arithmetic lives behind the C ABI (pse_shear_*, mirrors of PSEv1/SpecificShearFunction.h)."""
import ctypes

from . import system as _system
from ._lib import lib


class _cpp_function:
    """Owner of a pse_shear handle; exposes the reference's C++ method names (getShearRate, ...)."""

    def __init__(self, handle, keep=()):
        if not handle:
            raise RuntimeError("Error creating shear function")
        self._h, self._keep = handle, keep

    def getShearRate(self, timestep):
        return lib.pse_shear_rate(self._h, int(timestep) & 0xFFFFFFFF)

    def getStrain(self, timestep):
        return lib.pse_shear_strain(self._h, int(timestep) & 0xFFFFFFFF)

    def getOffset(self):
        return lib.pse_shear_offset(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.pse_shear_destroy(self._h)
            self._h = None


def _make(kind, args, offset, dt):
    arr = (ctypes.c_double * len(args))(*args) if args else None
    return _cpp_function(lib.pse_shear_create(kind, arr, len(args), int(offset), float(dt)))


class _shear_function:
    def __init__(self, zero="now"):
        self.cpp_function = None
        now = _system.current().getCurrentTimeStep()
        if zero == "now":
            self._offset = now
        else:
            if zero < 0:
                raise RuntimeError("Error creating shear function")  # negative zero (shear_function.py:20-22)
            if zero > now:
                raise RuntimeError("Error creating shear function")  # zero in the future (:23-25)
            self._offset = zero

    def get_shear_rate(self, timestep):
        return self.cpp_function.getShearRate(timestep)

    def get_strain(self, timestep):
        return self.cpp_function.getStrain(timestep)

    def get_offset(self):
        return self.cpp_function.getOffset()


class steady(_shear_function):
    def __init__(self, dt, shear_rate=0, zero="now"):
        _shear_function.__init__(self, zero)
        self.cpp_function = _make(1, [shear_rate], self._offset, dt)


class sine(_shear_function):
    def __init__(self, dt, shear_rate, shear_freq, zero="now"):
        if shear_rate <= 0:
            raise RuntimeError("Error creating shear function")
        if shear_freq <= 0:
            raise RuntimeError("Error creating shear function")
        _shear_function.__init__(self, zero)
        self.cpp_function = _make(2, [shear_rate, shear_freq], self._offset, dt)


class chirp(_shear_function):
    def __init__(self, dt, amplitude, omega_0, omega_f, periodT, zero="now"):
        _shear_function.__init__(self, zero)
        self.cpp_function = _make(3, [amplitude, omega_0, omega_f, periodT], self._offset, dt)


class tukey_window(_shear_function):
    def __init__(self, dt, periodT, tukey_param, zero="now"):
        if tukey_param <= 0 or tukey_param > 1:
            raise RuntimeError("Error creating Tukey window function")
        _shear_function.__init__(self, zero)
        self.cpp_function = _make(4, [periodT, tukey_param], self._offset, dt)


class windowed(_shear_function):
    def __init__(self, function_form, window):
        _shear_function.__init__(self, "now")
        h = lib.pse_shear_create_windowed(function_form.cpp_function._h, window.cpp_function._h)
        self.cpp_function = _cpp_function(h, keep=(function_form, window))

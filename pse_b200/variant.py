"""Wrapped-strain variant for the box tilt — PSEv1/variant.py:15-32 over
VariantShearFunction (PSEv1/VariantShearFunction.cc:17-43).  The reference's constructor calls an undefined
`_variant`; the signature is kept and the NameError fixed."""
from ._lib import lib


class _cpp_variant:
    def __init__(self, function_form, total_timestep, vmin, vmax):
        self._f, self._total, self._min, self._max = function_form, int(total_timestep), float(vmin), float(vmax)

    def getValue(self, timestep):
        return lib.pse_shear_variant_value(self._f.cpp_function._h, self._total, self._min, self._max, int(timestep) & 0xFFFFFFFF)


class shear_variant:
    def __init__(self, function_form, total_timestep, max_strain=0.5):
        if total_timestep <= 0:
            raise RuntimeError("Error creating variant")
        self.cpp_variant = _cpp_variant(function_form, total_timestep, -max_strain, max_strain)

    def get_value(self, timestep):
        return self.cpp_variant.getValue(timestep)

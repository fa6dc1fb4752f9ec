"""pyves_b200 -- B200-native biased-Langevin VES hot path behind the API of apallath/pyves.

The compute path is the CUDA library ``csrc/libpyves_b200.so`` (C ABI in
``include/pyves_b200.h``), reached through ``pyves_b200._lib``; the modules ``basis``, ``bias``,
``target``, ``potentials``, ``langevin_dynamics`` and ``config_creation`` mirror the reference's
``ves.*`` modules.  There is no CPU fallback.
"""
__version__ = "0.1.0"


def install_as_ves():
    """Register this package under the name ``ves`` so reference scripts (``from ves.basis import ...``)
    import the B200 implementation unchanged."""
    import importlib
    import sys
    me = sys.modules[__name__]
    sys.modules["ves"] = me
    for name in ("basis", "bias", "target", "potentials", "langevin_dynamics", "config_creation", "utils", "fes"):
        sys.modules["ves." + name] = importlib.import_module(__name__ + "." + name)
    return me

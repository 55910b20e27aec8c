"""faunus_b200 — B200-native (sm_100a) energy-evaluation hot path of Faunus behind its
EnergyTerm / Hamiltonian API. The native library (``libfaunus_b200.so``: CUDA kernels + C ABI +
C++ host adaptor layer) is required; there is no CPU fallback. See DESIGN.md."""
from . import config  # noqa: F401

__all__ = ["config", "native"]


def __getattr__(name):
    if name == "native":
        import importlib
        return importlib.import_module(".native", __name__)
    raise AttributeError(name)

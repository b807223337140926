"""qiskit-aer_b200: a B200-native statevector engine behind Qiskit Aer's QubitVector surface.

The directory name carries a hyphen (task layout); import it as ``qiskit_aer_b200``
through the shim module at the repo root.
"""
from .capi import B200Error, LIB_PATH  # noqa: F401
from .qubitvector import QubitVectorB200, rng_uniform  # noqa: F401

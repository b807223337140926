"""GPU: the reference's own held vectors (ref_1q_clifford / ref_2q_clifford / ref_non_clifford / ref_measure /
ref_reset / ref_rotation, e.g. ref_non_clifford.py:418-457,948-991) through the reference Controller on the B200
statevector, with and without Aer's fusion, in double and single precision."""
import pytest

import ref_vectors

pytestmark = pytest.mark.gpu


def _be():
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import aer_backend
    if not aer_backend.available():
        pytest.skip("Aer integration module not built (needs the reference tree at build time)")
    return aer_backend


@pytest.mark.parametrize("fusion", [False, True])
def test_b200_reproduces_the_reference_held_vectors(fusion):
    be = _be()
    for case in ref_vectors.load_cases():
        ref_vectors.check_case(be, case, "GPU", fusion=fusion, fusion_threshold=1)

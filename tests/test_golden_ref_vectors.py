"""CPU: the lifted reference vectors reproduce on the reference's own CPU path (pins the lifting script and the
driver conventions); the GPU twin is tests/test_gpu_ref_vectors.py."""
import pytest

import ref_vectors


def _be():
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import aer_backend
    if not aer_backend.available():
        pytest.skip("Aer integration module not built (needs the reference tree at build time)")
    return aer_backend


def test_fixture_is_populated():
    cases = ref_vectors.load_cases()
    assert len(cases) >= 100
    assert {c["file"] for c in cases} >= {"ref_1q_clifford", "ref_2q_clifford", "ref_non_clifford", "ref_measure", "ref_reset"}


@pytest.mark.parametrize("fusion", [False, True])
def test_reference_cpu_reproduces_its_held_vectors(fusion):
    be = _be()
    for case in ref_vectors.load_cases():
        ref_vectors.check_case(be, case, "CPU", fusion=fusion, fusion_threshold=1)

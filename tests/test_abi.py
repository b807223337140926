"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and
exports every symbol include/b200sv.h declares; calls fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200sv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import qiskit_aer_b200 as q
    lib = ctypes.CDLL(q.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), name
    # the ctypes binding covers the same set
    from qiskit_aer_b200 import capi
    assert set(capi.SIGNATURES) | {"b200sv_last_error"} == set(names)


def test_host_rng_matches_std_mt19937_64():
    import qiskit_aer_b200 as q
    r = q.rng_uniform(42, 4)
    # std::mt19937_64(42) first outputs / 2^64 (libstdc++ generate_canonical<double,53>)
    first = 13930160852258120406 / 2.0 ** 64
    assert abs(r[0] - first) < 1e-15
    assert all(0.0 <= x < 1.0 for x in r)


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import qiskit_aer_b200 as q
    with pytest.raises(q.B200Error, match="No CUDA device"):
        q.QubitVectorB200(4)

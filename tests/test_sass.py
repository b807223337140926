"""Static checks on the compiled kernels (cuobjdump; no GPU).  The tile kernels read gate matrices from the constant
bank through uniform registers (DFMA R, R, UR, R): that is what keeps a 2-qubit gate at 0.125 matrix fetches per DFMA.
ptxas silently falls back to vector-register operands (and local-memory copies of the parameter block) when a kernel
outgrows the uniform register file or indexes the parameter block per thread, so the property is pinned here."""
import os
import re
import shutil
import subprocess

import pytest

import qiskit_aer_b200  # noqa: F401
from qiskit_aer_b200 import capi

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name:
            kernels[name].append(line)
    return kernels


def test_tile_kernels_take_their_matrices_from_uniform_registers():
    kernels = _sass()
    tile = {k: v for k, v in kernels.items() if "tile_p" in k}
    assert len(tile) >= 6, sorted(tile)
    for name, lines in tile.items():
        fma = [l for l in lines if re.search(r"\b(DFMA|FFMA|FFMA2)\b", l)]
        assert len(fma) >= 256, (name, len(fma))
        if "tile_pipe2_kernelILi3E" in name:  # single-precision rounds: packed FP32 FMAs (Blackwell), no scalar FFMA left
            assert sum("FFMA2" in l for l in fma) >= 1024 and not any(re.search(r"\bFFMA\b", l) for l in fma), name
        vec = [l for l in fma if "UR" not in l]
        assert not vec, "%s: %d of %d FMAs lost their uniform-register operand" % (name, len(vec), len(fma))
        assert not any(re.search(r"\b(STL|LDL)\b", l) for l in lines if "DFMA" in l)


def test_tensor_path_and_async_copy_instructions_are_present():
    kernels = _sass()
    dmma = [k for k in kernels if "dense5_dmma" in k]
    assert dmma and all(sum("DMMA" in l for l in kernels[k]) == 128 for k in dmma)
    pipe2 = [k for k in kernels if "tile_pipe2" in k]
    assert pipe2 and all(any("LDGSTS" in l for l in kernels[k]) and any("SYNCS" in l for l in kernels[k]) for k in pipe2)

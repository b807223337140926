"""Batched-shot executor (config 5): per-shot Pauli noise riding on the tile passes.

* exact parity: the same per-shot Pauli realisation replayed shot by shot on the CPU oracle;
* statistical parity: means against the reference Controller's own noisy simulation (CPU, its own RNG)
  within 5 standard errors;
* the reference's per-shot noisy path on the B200 vector (Aer integration): identical counts for identical
  seeds (noise sampling, measurement and RNG all stay the reference's)."""
import numpy as np
import pytest

import opgen
from oracle.oracle import OracleQV

pytestmark = pytest.mark.gpu


def _ref_controller_subprocess(**kw):
    """The oracle's controller module and the Aer integration module are two builds of the same pybind
    module name; keep them in separate processes."""
    import os
    import pickle
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in.pkl"), "wb") as f:
            pickle.dump(kw, f)
        code = ("import pickle,sys;sys.path.insert(0,%r);import qiskit_aer_b200;from oracle import ref_controller as rc;"
                "kw=pickle.load(open(%r,'rb'));r=rc.run_circuit(**kw);"
                "pickle.dump({k:v for k,v in r['data'].items() if k.startswith('ev') or k=='counts'},open(%r,'wb'))"
                % (root, os.path.join(d, "in.pkl"), os.path.join(d, "out.pkl")))
        subprocess.check_call([sys.executable, "-c", code])
        return pickle.load(open(os.path.join(d, "out.pkl"), "rb"))


def _mats(op):
    from qiskit_aer_b200 import executor, fusion
    return list(op[2]), executor.colmajor(fusion.gate_matrix(op[1], op[3]))


@pytest.mark.parametrize("n,shots,batch", [(12, 24, 8), (13, 10, 16), (6, 12, 5)])
def test_shared_noise_realisation_matches_per_shot_oracle(n, shots, batch):
    from qiskit_aer_b200 import batched, circuits
    ops = circuits.random_noisy_circuit(n, 3, seed=n)
    rng = np.random.default_rng(4)
    nslots = sum(len(op[2]) for op in ops)
    codes = (rng.random((nslots, shots)) < 0.15) * rng.integers(1, 4, size=(nslots, shots))
    codes = codes.astype(np.uint8)
    obs = [([0], "Z"), ([1, 2], "ZZ"), ([n - 1, 0], "XY")]
    run = batched.BatchedShotsRunner(n, batch)
    out = run.run(ops, shots, seed=1, observables=obs, measure=False, codes=codes)
    want = np.zeros(len(obs))
    last_state = None
    for s in range(shots):
        o = OracleQV(n)
        slot = 0
        for op in ops:
            qs, m = _mats(op)
            o.apply_matrix(qs, m)
            for q in qs:
                c = int(codes[slot, s])
                slot += 1
                if c:
                    o.apply_pauli([q], "IXYZ"[c])
        want += [o.expval_pauli(qs, pl) for qs, pl in obs]
        last_state = o.vector()
    np.testing.assert_allclose(out["expval"], want / shots, atol=1e-10)
    # the container still holds the last batch: its last used state is the last shot's state
    nb_last = shots - (shots - 1) // batch * batch
    got = run.qv.vector(offset=(nb_last - 1) << n, count=1 << n)
    assert opgen.fidelity_gap(last_state, got) < 1e-10
    if n >= 12:
        assert out["passes"] < len(ops) * ((shots + batch - 1) // batch) // 3  # gates + noise share passes
    run.close()


def test_noisy_means_agree_with_reference_controller():
    from oracle import ref_controller
    from qiskit_aer_b200 import batched, circuits, noise
    if not ref_controller.available():
        pytest.skip("oracle/_ref/controller_wrappers.so not built")
    n, shots, p1, p2 = 12, 4000, 0.02, 0.08
    ops = circuits.random_noisy_circuit(n, 2, seed=7)
    obs = [([q], "Z") for q in range(0, n, 3)] + [([0, 1], "ZZ"), ([2], "X"), ([5, 7], "YY")]
    run = batched.BatchedShotsRunner(n, 1024)
    out = run.run(ops, shots, seed=11, p1=p1, p2=p2, observables=obs, measure=False)
    ideal = run.run(ops, 1, seed=0, observables=obs, measure=False)
    ref = _ref_controller_subprocess(n=n, ops=ops, shots=shots, seed=5, fusion=False, expvals=obs,
                                     noise_model=noise.noise_model_dict(p1, p2), measure=False)
    ref_ev = np.array([ref["ev%d" % i] for i in range(len(obs))])
    se = np.sqrt(out["expval_stderr"] ** 2 + 1.0 / shots)  # reference's own sampling error <= 1/sqrt(shots)
    assert np.all(np.abs(out["expval"] - ref_ev) < 5 * se + 1e-3), (out["expval"], ref_ev)
    # and the noise must actually have done something
    assert np.max(np.abs(ideal["expval"] - out["expval"])) > 0.02
    run.close()


def test_reference_noisy_per_shot_path_on_b200_vector():
    from qiskit_aer_b200 import aer_backend, circuits, noise
    if not aer_backend.available():
        pytest.skip("Aer integration module not built")
    n, shots = 8, 300
    ops = circuits.random_noisy_circuit(n, 3, seed=2)
    nm = noise.noise_model_dict(0.02, 0.08)
    kw = dict(shots=shots, seed=17, fusion=False, noise_model=nm)
    gpu = aer_backend.run_circuit(n, ops, device="GPU", **kw)
    cpu = aer_backend.run_circuit(n, ops, device="CPU", **kw)
    assert gpu["metadata"]["device"] == "GPU"
    assert gpu["data"]["counts"] == cpu["data"]["counts"]

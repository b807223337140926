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


@pytest.mark.parametrize("n,shots,batch", [(12, 24, 8), (13, 10, 16), (6, 12, 5), (20, 5, 4)])  # 20: the config-5 width
def test_shared_noise_realisation_matches_per_shot_oracle(n, shots, batch):
    from qiskit_aer_b200 import batched, circuits
    ops = circuits.random_noisy_circuit(n, 3 if n < 20 else 6, seed=n)
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


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 5e-6)])
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_per_state_matrices_in_one_launch(dtype, tol, k):
    """b200sv_apply_batched_matrix: every state applies its own (scaled) matrix of a table, or none."""
    import qiskit_aer_b200 as q
    n, S, nm = 7, 37, 5
    rng = np.random.default_rng(40 + k)
    mats = [opgen.colmajor(opgen.haar_unitary(rng, 1 << k) * rng.uniform(0.5, 1.5)) for _ in range(nm)]
    index = rng.integers(-1, nm, size=S)
    scale = rng.uniform(0.5, 2.0, size=S)
    qubits = opgen.pick(rng, n, k)
    gpu = q.QubitVectorB200(n, dtype, num_states=S)
    states = [opgen.random_state(rng, n, dtype) for _ in range(S)]
    gpu.set_state(np.concatenate(states))
    gpu.apply_batched_matrix(qubits, np.concatenate(mats), index, scale)
    got = gpu.vector().reshape(S, -1)
    for s in range(S):
        ora = OracleQV(n)
        ora.set_state(states[s].astype(np.complex128))
        if index[s] >= 0:
            ora.apply_matrix(qubits, mats[index[s]] * scale[s])
        assert np.max(np.abs(got[s] - ora.vector())) < tol * 4, (s, index[s])


@pytest.mark.parametrize("nq,S", [(5, 10), (6, 11), (7, 3), (5, 100)])
def test_five_qubit_block_on_a_batch_whose_group_count_is_not_a_multiple_of_8(nq, S):
    """The DMMA kernel walks batches of 8 groups: containers with a ragged group count take the register kernel."""
    import qiskit_aer_b200 as q
    rng = np.random.default_rng(nq * 100 + S)
    U = opgen.colmajor(opgen.haar_unitary(rng, 32))
    qubits = opgen.pick(rng, nq, 5)
    gpu = q.QubitVectorB200(nq, np.complex128, num_states=S)
    states = [opgen.random_state(rng, nq) for _ in range(S)]
    gpu.set_state(np.concatenate(states))
    gpu.apply_matrix(qubits, U)
    got = gpu.vector().reshape(S, -1)
    for s in range(S):
        ora = OracleQV(nq)
        ora.set_state(states[s])
        ora.apply_matrix(qubits, U)
        assert np.max(np.abs(got[s] - ora.vector())) < 1e-12, s


@pytest.mark.parametrize("seed", [0, 1])
def test_noisy_pass_slot_forms(seed):
    """Slot rounds of noisy passes on the device: bare cx gates (either control) folded into the load offsets,
    1-qubit gates by kind (dense / diagonal / real / unpaired), sampled Paulis before and between them -- against the
    oracle replaying every state's own Pauli draws."""
    import qiskit_aer_b200 as q
    n, S = 14, 5
    rng = np.random.default_rng(300 + seed)
    cx_first = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)
    cx_second = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    ops, nslots = [], 0
    for layer in range(6):
        for qb in range(n):
            kind = int(rng.integers(4))
            if kind == 1:
                u = np.diag(np.exp(1j * rng.uniform(0, 6.28, 2)))
            elif kind == 2:
                th = rng.uniform(0, 6.28)
                u = np.array([[1, 1], [1, -1]]) / np.sqrt(2) if rng.random() < 0.5 else np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            else:
                u = opgen.haar_unitary(rng, 2)
            if layer == 4 and qb % 3 == 0:
                continue
            ops.append(("dense", [qb], opgen.colmajor(np.asarray(u, dtype=np.complex128))))
            ops.append(("pauli", qb, nslots)); nslots += 1
        perm = rng.permutation(n)
        for i in range(n // 2):
            a, b = int(perm[2 * i]), int(perm[2 * i + 1])
            ops.append(("dense", [a, b], opgen.colmajor(cx_first if rng.random() < 0.5 else cx_second)))
            for qb in (a, b):
                ops.append(("pauli", qb, nslots)); nslots += 1
    codes = rng.choice(4, size=(nslots, S), p=[0.8, 0.07, 0.07, 0.06]).astype(np.uint8)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    qv = q.QubitVectorB200(n, np.complex128, num_states=S)
    qv.initialize_from_vector(np.concatenate(states))
    ngates = sum(1 for o in ops if o[0] == "dense")
    passes = qv.apply_op_sequence(ops, codes)
    assert passes <= ngates // 16
    got = qv.vector().reshape(S, -1)
    for si, st in enumerate(states):
        o = OracleQV(n)
        o.set_state(st)
        for op in ops:
            if op[0] == "dense":
                o.apply_matrix(op[1], op[2])
            elif codes[op[2], si]:
                o.apply_pauli([op[1]], "IXYZ"[int(codes[op[2], si])])
        assert np.max(np.abs(got[si] - o.vector())) < 1e-12
    qv.close()

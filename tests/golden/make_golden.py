"""Generates the committed golden fixtures by running the UNMODIFIED reference
in the dev container (needs /root/reference-built oracle/_ref/*.so):

    python tests/golden/make_golden.py

* kernel_n{N}_{dtype}.npz  -- seeded random op streams replayed on the reference's
  QubitVector (oracle/_ref/libaer_qv_ref.so): initial state, per-op serialised
  arguments, final state, norm/probabilities/expval/sample_measure answers.
* circuit_qv{N}.npz        -- Quantum-Volume style circuits run through the
  reference Controller (oracle/_ref/controller_wrappers.so; Fusion on/off):
  final statevector, Pauli expectation values and sampled counts for a fixed
  seed_simulator.

Fixtures are small (n <= 12) so they live in git; the GPU box never needs the
reference tree.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import opgen  # noqa: E402
from oracle import ref_controller  # noqa: E402
from oracle.oracle import RefQV  # noqa: E402


def enc(x):
    if isinstance(x, np.ndarray):
        if np.iscomplexobj(x):
            return {"c": np.stack([x.real, x.imag], -1).tolist()}
        return x.tolist()
    if isinstance(x, complex):
        return {"z": [x.real, x.imag]}
    if isinstance(x, (list, tuple)):
        return [enc(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    return x


def kernel_fixture(n, dtype, seed, count):
    float32 = dtype == np.complex64
    kinds = None
    if float32:  # the reference's float diagonal kernel is unsafe (see tests/test_oracle_pin.py)
        kinds = ["matrix", "mcx", "mcy", "mcswap", "mcphase", "mcu", "pauli", "multiplexer", "permutation"]
    rng = np.random.default_rng(seed)
    psi0 = opgen.random_state(rng, n, dtype)
    ops = opgen.random_ops(seed + 1, n, count, kinds=kinds)
    ref = RefQV(n, dtype)
    ref.set_state(psi0)
    for op in ops:
        opgen.apply(ref, op)
    final = ref.vector()
    paulis = opgen.random_paulis(seed + 2, n, 12)
    expv = np.array([ref.expval_pauli(q, p) for q, p in paulis])
    pq = [opgen.pick(rng, n, k) for k in (1, 2, 3, min(5, n))]
    probs = [ref.probabilities(q) for q in pq]
    rnds = rng.random(64)
    samples = ref.sample_measure(rnds)
    kq = opgen.pick(rng, n, 2)
    K = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    meta = {"n": n, "ops": enc(ops), "paulis": enc(paulis), "prob_qubits": pq, "kraus_qubits": kq}
    name = "kernel_n%d_%s.npz" % (n, "f32" if float32 else "f64")
    np.savez_compressed(os.path.join(HERE, name), meta=json.dumps(meta), psi0=psi0, final=final, expval=expv,
                        norm=ref.norm(), rnds=rnds, samples=samples, kraus=opgen.colmajor(K),
                        kraus_norm=ref.norm(kq, opgen.colmajor(K)),
                        **{"probs%d" % i: p for i, p in enumerate(probs)})
    print("wrote", name)


def qv_ops(n, depth, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(depth):
        perm = rng.permutation(n)
        for i in range(n // 2):
            ops.append(("unitary", [int(perm[2 * i]), int(perm[2 * i + 1])], opgen.haar_unitary(rng, 4)))
    return ops


def circuit_fixture(n, depth, seed, shots):
    ops = qv_ops(n, depth, seed)
    paulis = opgen.random_paulis(seed + 5, n, 8, max_weight=4)
    out = {}
    for tag, fusion in (("fused", True), ("plain", False)):
        r = ref_controller.run_circuit(n, ops, shots=shots, seed=seed, fusion=fusion, fusion_threshold=1,
                                       expvals=paulis, save_statevector=True)
        out["sv_" + tag] = np.asarray(r["data"]["sv"])
        out["ev_" + tag] = np.array([r["data"]["ev%d" % i] for i in range(len(paulis))])
        counts = np.zeros(1 << n, dtype=np.int64)
        for k, v in r["data"]["counts"].items():
            counts[int(k, 16)] = v
        out["counts_" + tag] = counts
    meta = {"n": n, "depth": depth, "seed": seed, "shots": shots, "paulis": enc(paulis),
            "qubits": [op[1] for op in ops]}
    name = "circuit_qv%d.npz" % n
    np.savez_compressed(os.path.join(HERE, name), meta=json.dumps(meta),
                        unitaries=np.stack([op[2] for op in ops]), **out)
    print("wrote", name)


if __name__ == "__main__":
    kernel_fixture(6, np.complex128, 11, 80)
    kernel_fixture(10, np.complex128, 12, 80)
    kernel_fixture(12, np.complex128, 13, 60)
    kernel_fixture(10, np.complex64, 14, 60)
    circuit_fixture(10, 6, 77, 2000)
    circuit_fixture(12, 4, 78, 2000)

"""Config 5: noisy 20-qubit random circuit, depolarizing noise, 10k shots, Pauli expectation values.

    python tools/bench_noisy.py [--qubits 20 --depth 20 --shots 10000 --batch 2048]

Times (a) the batched-shot executor (sampled Paulis riding on tile passes), (b) the reference's per-shot
noisy path running on the B200 vector through the Aer integration module, (c) the reference CPU path on
a bounded number of shots.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qiskit_aer_b200  # noqa: E402,F401
from qiskit_aer_b200 import aer_backend, batched, circuits, noise  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=20)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--shots", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--p1", type=float, default=1e-3)
    ap.add_argument("--p2", type=float, default=1e-2)
    ap.add_argument("--aer-shots", type=int, default=500)
    ap.add_argument("--cpu-shots", type=int, default=200)
    a = ap.parse_args()
    n = a.qubits
    ops = circuits.random_noisy_circuit(n, a.depth, seed=1)
    rng = np.random.default_rng(0)
    obs = []
    for _ in range(10):  # ~10 random weight<=4 Pauli strings (EstimatorV2 lowering: one save_expval each)
        k = int(rng.integers(1, 5))
        qs = [int(q) for q in rng.choice(n, size=k, replace=False)]
        obs.append((qs, "".join("XYZ"[int(c)] for c in rng.integers(0, 3, size=k))))
    out = {"config": {"workload": "noisy_random_circuit", "qubits": n, "depth": a.depth, "gates": len(ops),
                      "shots": a.shots, "p1": a.p1, "p2": a.p2, "observables": len(obs)}}
    import torch
    run = batched.BatchedShotsRunner(n, a.batch)
    run.run(ops, a.batch, seed=1, p1=a.p1, p2=a.p2, observables=obs, measure=True)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = run.run(ops, a.shots, seed=2, p1=a.p1, p2=a.p2, observables=obs, measure=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["batched"] = {"seconds": dt, "shots_per_s": a.shots / dt, "batch_states": a.batch,
                      "hbm_passes": res["passes"], "expval": [float(x) for x in res["expval"]],
                      "stderr": [float(x) for x in res["expval_stderr"]]}
    run.close()
    nm = noise.noise_model_dict(a.p1, a.p2)
    if aer_backend.available():
        kw = dict(seed=3, fusion=False, noise_model=nm, expvals=obs)
        aer_backend.run_circuit(n, ops, device="GPU", shots=16, **kw)
        t0 = time.perf_counter()
        r = aer_backend.run_circuit(n, ops, device="GPU", shots=a.aer_shots, **kw)
        dt = time.perf_counter() - t0
        out["aer_per_shot_on_b200"] = {"seconds": dt, "shots": a.aer_shots, "shots_per_s": a.aer_shots / dt,
                                       "expval": [float(r["data"]["ev%d" % i]) for i in range(len(obs))]}
        # the reference's BatchShotsExecutor (batched_shots_gpu=True) driving the B200 container
        bkw = dict(kw, batched_shots_gpu=True, batched_shots_gpu_max_qubits=max(n, 16))
        aer_backend.run_circuit(n, ops, device="GPU", shots=64, **bkw)
        t0 = time.perf_counter()
        r = aer_backend.run_circuit(n, ops, device="GPU", shots=a.shots, **bkw)
        dt = time.perf_counter() - t0
        out["aer_batched_shots_gpu_on_b200"] = {"seconds": dt, "shots": a.shots, "shots_per_s": a.shots / dt,
                                                "batched": r["metadata"].get("batched_shots_optimization"),
                                                "expval": [float(r["data"]["ev%d" % i]) for i in range(len(obs))]}
        t0 = time.perf_counter()
        r = aer_backend.run_circuit(n, ops, device="CPU", shots=a.cpu_shots, **kw)
        dt = time.perf_counter() - t0
        out["reference_cpu"] = {"seconds": dt, "shots": a.cpu_shots, "shots_per_s": a.cpu_shots / dt,
                                "cores": os.cpu_count(),
                                "expval": [float(r["data"]["ev%d" % i]) for i in range(len(obs))]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

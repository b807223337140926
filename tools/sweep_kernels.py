"""Per-kernel timing sweep on one B200 (CUDA events, warm, state larger than L2).

    python tools/sweep_kernels.py --qubits 30 > gpurun_out/sweep.json

Reports achieved algorithmic GB/s (BASELINE.md section 3 byte counts) per kernel and qubit placement.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qiskit_aer_b200 as q  # noqa: E402
from qiskit_aer_b200 import circuits  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dtype", default="f64")
    args = ap.parse_args()
    n = args.qubits
    dt = np.complex128 if args.dtype == "f64" else np.complex64
    S = 16 if args.dtype == "f64" else 8
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        buf = torch.empty((1 << n) * 2, dtype=torch.float64 if args.dtype == "f64" else torch.float32, device=dev)
    qv = q.QubitVectorB200(n, dt, external_ptr=buf.data_ptr(), stream=stream.cuda_stream)
    qv.initialize()
    rng = np.random.default_rng(0)
    # spread the state
    for i in range(0, n - 1, 2):
        qv.apply_matrix([i, i + 1], circuits.haar_unitary(rng, 4).reshape(-1, order="F"))
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    rows = []

    def timeit(name, fn, bytes_alg):
        fn()
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        gbs = bytes_alg / (ms / 1e3) / 1e9
        rows.append({"kernel": name, "ms": ms, "GBps": gbs, "frac_measured_peak": gbs / peak})
        print("%-34s %9.3f ms %9.1f GB/s  %.3f" % (name, ms, gbs, gbs / peak), file=sys.stderr, flush=True)

    full = 2 * S * 2.0 ** n
    places = {"low": lambda k: list(range(k)), "high": lambda k: list(range(n - k, n)),
              "mid": lambda k: [7, 12, 19, 23, 10][:k], "mixed": lambda k: [0, n - 1, 3, 17, 9][:k]}
    for k in (1, 2, 3, 4, 5):
        U = circuits.haar_unitary(rng, 1 << k).reshape(-1, order="F")
        for pname, pf in places.items():
            qs = pf(k)
            timeit("dense_k%d_%s" % (k, pname), lambda: qv.apply_matrix(qs, U), full)
    for k in (1, 3, 5, 8, 10, 12, 14, 16):
        d = np.exp(2j * np.pi * rng.random(1 << k))
        for pname in ("low", "high"):
            qs = places[pname](k)
            timeit("diag_k%d_%s" % (k, pname), lambda: qv.apply_diagonal_matrix(qs, d), full)
    for pname, qs in (("t_low", [0]), ("t_high", [n - 1]), ("c1_t_low", [n - 1, 0]), ("c1_t_high", [0, n - 1]),
                      ("c2", [5, 20, 11])):
        c = len(qs) - 1
        timeit("mcx_%s" % pname, lambda: qv.apply_mcx(qs), full / 2 ** c)
    timeit("mcswap_c0", lambda: qv.apply_mcswap([3, n - 2]), full / 2)
    timeit("mcphase_2q", lambda: qv.apply_mcphase([4, n - 3], np.exp(0.3j)), full / 4)
    U2 = circuits.haar_unitary(rng, 2).reshape(-1, order="F")
    timeit("mcu_c1", lambda: qv.apply_mcu([6, n - 1], U2), full / 2)
    timeit("pauli_XYZ", lambda: qv.apply_pauli([1, n - 1, 9], "XYZ"), full)
    timeit("pauli_ZZ", lambda: qv.apply_pauli([1, n - 1], "ZZ"), full)
    half = full / 2
    timeit("norm", lambda: qv.norm(), half)
    timeit("expval_ZZ", lambda: qv.expval_pauli([2, n - 1], "ZZ"), half)
    timeit("expval_XY", lambda: qv.expval_pauli([2, n - 1], "XY"), half)
    timeit("probabilities_3q", lambda: qv.probabilities([0, 13, n - 1]), half)
    timeit("probabilities_10q", lambda: qv.probabilities(list(range(3, 13))), half)
    K = (rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))).reshape(-1, order="F")
    timeit("norm_matrix_2q", lambda: qv.norm([3, n - 1], K), half)
    r = rng.random(1024)
    timeit("sample_measure_1024", lambda: qv.sample_measure(r), half)
    print(json.dumps({"qubits": n, "dtype": args.dtype, "peak_GBps": peak, "rows": rows}))


if __name__ == "__main__":
    main()

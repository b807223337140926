"""Batched-shot executor: many noisy-shot statevectors packed into one container, one launch per pass.

B200-native take on `BatchShotsExecutor` (/root/reference/src/simulators/batch_shots_executor.hpp:
run_circuit_shots :317, apply_ops_batched_shots_for_group :500, apply_batched_expval :731) for circuits
whose noise is a Pauli mixture sampled per shot (`sample_noise_at_runtime`, noise_model.hpp:320): the
sampled Paulis do not get their own launches (qubitvector_thrust.hpp:2892) -- they ride, together with
5-10 gates, on the tile-blocked passes of b200sv_apply_op_sequence, selected per state by a code table.
Observables are reduced per state in one launch per Pauli string (batched_expval_pauli :2683) and averaged
over shots on the host, as `save_expval` does (state.hpp:441-446); final measurements draw one sample per
state."""
import numpy as np

from .executor import _small_dense, apply_op
from .noise import sample_pauli_codes
from .qubitvector import QubitVectorB200


class BatchedShotsRunner:
    def __init__(self, num_qubits, batch_states, device=0):
        self.n = int(num_qubits)
        self.batch = int(batch_states)
        self.qv = QubitVectorB200(self.n, np.complex128, num_states=self.batch, device=device)
        self.passes = 0

    def close(self):
        self.qv.close()

    def _lower(self, ops, noisy, p1, p2, slots):
        """Circuit -> op sequence for the C ABI: dense gates plus per-state Pauli ops after noisy gates."""
        seq, occ = [], 0
        for op in ops:
            g = _small_dense(op)
            if g is None and op[0] == "gate" and op[1] in ("cx", "rz"):
                from .fusion import gate_matrix
                from .executor import colmajor
                g = (list(op[2]), colmajor(gate_matrix(op[1], op[3])))
            if g is None:
                raise ValueError("batched executor: unsupported op %r" % (op[:2],))
            seq.append(("dense", g[0], g[1]))
            if noisy(op):
                for q, sl in zip(g[0], slots[occ]):
                    seq.append(("pauli", q, sl))
                occ += 1
        return seq

    def run(self, ops, shots, seed, p1=0.0, p2=0.0, observables=(), measure=True, codes=None):
        """Returns {"expval": [...], "expval_stderr": [...], "samples": uint64[shots], "passes": n}.

        `codes` (optional) injects the per-shot Pauli samples ([nslots][shots]) instead of drawing them --
        used by the parity tests to share the exact noise realisation with the per-shot oracle."""
        rng = np.random.default_rng(seed)
        noisy_qubits = [tuple(op[1] if op[0] == "unitary" else op[2]) for op in ops]
        is_noisy = (lambda op: True) if (p1 > 0 or p2 > 0 or codes is not None) else (lambda op: False)
        if codes is None and (p1 > 0 or p2 > 0):
            codes, slots = sample_pauli_codes(rng, noisy_qubits, shots, p1, p2)
        elif codes is not None:
            slots, s = [], 0
            for qs in noisy_qubits:
                slots.append(tuple(range(s, s + len(qs))))
                s += len(qs)
        else:
            slots = [()] * len(ops)
        seq = self._lower(ops, is_noisy, p1, p2, slots)
        ev_sum = np.zeros(len(observables))
        ev_sq = np.zeros(len(observables))
        samples = np.zeros(shots, dtype=np.uint64)
        self.passes = 0
        for b0 in range(0, shots, self.batch):
            nb = min(self.batch, shots - b0)
            self.qv.initialize()
            bc = None
            if codes is not None:
                bc = np.zeros((codes.shape[0], self.batch), dtype=np.uint8)
                bc[:, :nb] = codes[:, b0:b0 + nb]
            self.passes += self.qv.apply_op_sequence(seq, bc)
            for i, (qs, pl) in enumerate(observables):
                v = np.atleast_1d(self.qv.expval_pauli(qs, pl))[:nb]
                ev_sum[i] += v.sum()
                ev_sq[i] += (v * v).sum()
            if measure:
                r = rng.random((self.batch, 1))
                samples[b0:b0 + nb] = np.atleast_2d(self.qv.sample_measure(r))[:nb, 0]
        mean = ev_sum / shots
        var = np.maximum(ev_sq / shots - mean ** 2, 0.0)
        return {"expval": mean, "expval_stderr": np.sqrt(var / shots), "samples": samples, "passes": self.passes}

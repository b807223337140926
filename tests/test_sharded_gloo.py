"""world_size-2/4 `gloo` tests (CPU) of the sharded-state host logic: the swap planner (logical ->
physical map, Belady eviction, no swap-back), the sliced pairwise exchange (both the contiguous-run
and the pack/unpack path), chunk-index resolution of global diagonal qubits, and the cross-rank
sampler / expectation values.  The chunk arithmetic is the CPU oracle (tests/cpu_chunk.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, n, min_run_bits, slice_amps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import opgen
        import qiskit_aer_b200  # noqa: F401
        from cpu_chunk import CpuChunk
        from oracle.oracle import OracleQV
        from qiskit_aer_b200 import circuits, executor, fusion, sharded
        gb = int(np.log2(world))
        chunk = CpuChunk(n - gb)
        run = sharded.ShardedRunner(chunk, n, rank, world, None, chunk.buf, slice_amps=slice_amps,
                                    min_run_bits=min_run_bits)
        ops = circuits.quantum_volume(n, 4, seed=5) + circuits.qft(n)
        fused = fusion.fuse(ops, max_qubit=3)
        run.initialize()
        plan = run.plan(fused)
        nswaps = sum(1 for p in plan if p[0] in ("swap", "mswap"))
        for p in plan:
            run.apply(p)
        # reference: the same circuit on one unsharded oracle state
        ref = OracleQV(n)
        executor.apply_ops(ref, ops)
        # gather physical chunks, undo the qubit map
        mine = torch.from_numpy(chunk.ora.psi.copy().view(np.float64))
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        phys = np.concatenate([p.numpy().view(np.complex128) for p in parts])
        idx = np.arange(1 << n, dtype=np.uint64)
        pidx = np.zeros_like(idx)
        for lq, pp in enumerate(run.phys):
            pidx |= ((idx >> np.uint64(lq)) & np.uint64(1)) << np.uint64(pp)
        logical = phys[pidx.astype(np.int64)]
        gap = opgen.fidelity_gap(ref.vector(), logical)
        maxerr = float(np.max(np.abs(ref.vector() - logical)))
        # reductions
        nrm = run.norm()
        ev_err = 0.0
        for qs, pl in opgen.random_paulis(3, n, 8, max_weight=3):
            ev_err = max(ev_err, abs(run.expval_pauli(qs, pl) - ref.expval_pauli(qs, pl)))
        # (expval may have swapped qubits: recheck the state is still the same logical state)
        rn = np.random.default_rng(1).random(300)
        samples = run.sample_measure(rn)
        ok_samples = bool(np.array_equal(samples, ref.sample_measure(rn)))
        if rank == 0:
            q.put({"gap": gap, "maxerr": maxerr, "nswaps": nswaps, "norm": nrm, "ev_err": ev_err,
                   "samples": ok_samples})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,min_run_bits,slice_amps", [
    (2, 8, 3, 1 << 4),     # contiguous-run exchange, several slices
    (2, 7, 20, 1 << 3),    # forces the pack/unpack path (no position is 'high enough')
    (4, 9, 3, 1 << 20),    # two global qubits, single-slice transfers
])
def test_sharded_circuit_matches_unsharded(world, n, min_run_bits, slice_amps):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, min_run_bits, slice_amps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["nswaps"] > 0, "circuit should need global-qubit exchanges"
    assert res["gap"] < 1e-10 and res["maxerr"] < 1e-12, res
    assert abs(res["norm"] - 1.0) < 1e-12
    assert res["ev_err"] < 1e-10
    assert res["samples"], "sharded sampler must reproduce the unsharded sampled indices"

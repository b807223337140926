"""Run under torchrun on >= 2 GPUs: sharded NCCL path vs the CPU oracle (n small enough for the CPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded.py --qubits 20
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=20)
    ap.add_argument("--slice-amps", type=int, default=1 << 14)
    ap.add_argument("--min-run-bits", type=int, default=12)
    args = ap.parse_args()
    import opgen
    import qiskit_aer_b200 as q
    from oracle.oracle import OracleQV
    from qiskit_aer_b200 import circuits, executor, fusion, sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = args.qubits
    nl = n - int(np.log2(world))
    stream = torch.cuda.Stream(device=dev)
    qv = q.QubitVectorB200(nl, np.complex128, device=local)   # library-owned memory (IPC exportable)
    qv.set_stream(stream.cuda_stream)
    buf = qv.torch_view()
    ok = True
    runners = {}
    # NVLink peer-swap kernel over CUDA IPC, NCCL contiguous-run slices, NCCL pack/unpack slices
    for mode, min_run_bits in (("p2p", args.min_run_bits), ("nccl", args.min_run_bits), ("nccl", 40)):
        with torch.cuda.stream(stream):
            run = sharded.ShardedRunner(qv, n, rank, world, stream, buf, slice_amps=args.slice_amps,
                                        min_run_bits=min_run_bits, exchange=mode)
        runners[(mode, min_run_bits)] = run
        ops = circuits.quantum_volume(n, 5, seed=11) + circuits.qft(n)
        fused = fusion.fuse(ops, max_qubit=4)
        run.initialize()
        plan = run.plan(fused)
        for p in plan:
            run.apply(p)
        ref = OracleQV(n)
        executor.apply_ops(ref, ops)
        ev_err = max(abs(run.expval_pauli(qs, pl) - ref.expval_pauli(qs, pl))
                     for qs, pl in opgen.random_paulis(3, n, 6, max_weight=3))
        rn = q.rng_uniform(99, 500)
        same = np.array_equal(run.sample_measure(rn), ref.sample_measure(rn))
        # after restore_order the chunks are the plain slices of the logical state
        qv.synchronize()
        mine = qv.vector()
        want = ref.vector()[rank << nl:(rank + 1) << nl]
        err = float(np.max(np.abs(mine - want)))
        nsw = sum(1 for p in plan if p[0] in ("swap", "mswap"))
        good = err < 1e-12 and ev_err < 1e-10 and same and nsw > 0
        ok = ok and good
        print("rank %d %s min_run_bits=%d swaps=%d max|err|=%.2e ev_err=%.2e samples_equal=%s" %
              (rank, mode, min_run_bits, nsw, err, ev_err, same), flush=True)
    # ---- the path bench.py times at N > 1: the C++ sharded executor (csrc/sharded.cu), one shard per process --
    # un-fused gates -> epoch plan -> tile passes on the shard + staged / pipelined exchange (flag words over NVLink)
    del runners
    qv.close()
    torch.cuda.synchronize()
    for label, env, staging in (("cpp-staged", {"B200SV_SHARD_MIN_RUN_BITS": "5", "B200SV_SHARD_SLAB_BITS": "2"}, 1 << 22),
                                ("cpp-inplace", {}, 0)):
        os.environ.update(env)
        st = sharded.ShardedState(n, world=world, rank=rank, device=local, dist=dist, staging_bytes=staging)
        ops = circuits.quantum_volume(n, 6, seed=17) + circuits.qft(n)
        ref = OracleQV(n)
        executor.apply_ops(ref, ops)
        for rep in range(2):
            st.initialize()
            st.apply_ops(ops)
        stats = st.stats()
        ev_err = max(abs(st.expval_pauli(qs, pl) - ref.expval_pauli(qs, pl))
                     for qs, pl in opgen.random_paulis(5, n, 6, max_weight=3))
        rn = q.rng_uniform(77, 500)
        same = np.array_equal(st.sample_measure(rn), ref.sample_measure(rn))
        mine = st.shard_vector(rank)
        err = float(np.max(np.abs(mine - ref.vector()[rank << nl:(rank + 1) << nl])))
        good = err < 1e-12 and ev_err < 1e-10 and same and stats["exchanges"] > 0
        ok = ok and good
        print("rank %d %s exchanges=%d staged=%d inplace=%d overlapped=%d max|err|=%.2e ev_err=%.2e samples_equal=%s" %
              (rank, label, stats["exchanges"], stats["staged"], stats["inplace"], stats["overlapped_passes"], err, ev_err,
               same), flush=True)
        st.close()
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("SHARDED_CHECK " + ("OK" if t.item() == 0 else "FAILED"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the statevector amplitude-update path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 engine
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (AerSimulator CPU equivalent)

A "step" is one full pass of the workload circuit over a freshly initialised state:
Quantum Volume, depth 10, double precision, 33 qubits per GPU (128 GiB of amplitudes; weak scaling:
33 + log2(N) qubits on N GPUs, i.e. QV-36 on 8), gates fused by the engine's own fusion pass.
`value` = circuit-level amplitude updates per second (sum over the circuit's gates of the amplitudes
an un-fused pass would write, BASELINE.md section 3 -- independent of how an engine fuses) with the
state resident in HBM; `e2e` = the same through the public host API (circuit in host memory ->
fusion -> C-ABI calls -> sampled counts + Pauli expectation values back on the host).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU legs (reference arm, cpu_baseline) are meant to use every
# host core, and libgomp reads the variable when it is loaded -- so fix it before anything OpenMP-linked is imported.
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    if os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        os.environ["OMP_NUM_THREADS"] = str(host_cores())

import numpy as np  # noqa: E402

DEPTH = 10
SHOTS = 1024
AMP_BYTES = 16
CDTYPE = None  # numpy complex dtype of the amplitudes, set from --precision


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload(args, world):
    from qiskit_aer_b200 import circuits, fusion
    n = args.qubits if args.qubits else 33 + int(np.log2(world))
    if args.workload == "qft":
        n = args.qubits if args.qubits else 30 + int(np.log2(world))
        ops = circuits.qft(n)
        name = "qft%d_fused" % n
    else:
        ops = circuits.quantum_volume(n, args.depth, seed=1234)
        name = "qv%d_depth%d_fused" % (n, args.depth)
    return n, name, ops, circuits.amplitudes_written(ops, n)


# ------------------------------------------------------------------------------------------ reference arm
def reference_arm(args):
    """The reference's own CPU implementation (Controller + Fusion + OpenMP/AVX2 QubitVector, built
    unmodified into oracle/_ref/controller_wrappers.so) on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_controller
    from qiskit_aer_b200 import circuits
    world = args.gpus
    n_full, name, _, _ = workload(args, world)
    cores = host_cores()
    if not ref_controller.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/controller_wrappers.so not built"}))
        return
    used = {"threads": None}

    def run(n):
        ops = circuits.qft(n) if args.workload == "qft" else circuits.quantum_volume(n, args.depth, seed=1234)
        t0 = time.perf_counter()
        r = ref_controller.run_circuit(n, ops, shots=SHOTS, seed=1234, threads=cores, fusion=True,
                                       fusion_max_qubit=5, fusion_threshold=14,
                                       expvals=[([0, 1, n - 1], "ZXY")])
        used["threads"] = int(r["metadata"].get("parallel_state_update", 0)) or used["threads"]
        return time.perf_counter() - t0, float(r["time_taken"]), circuits.amplitudes_written(ops, n)

    # size the sample: ~6 s per step (time doubles per qubit), capped by host memory
    n = 22
    t, _, _ = run(n)
    t, _, _ = run(n)
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 32 << 30
    while n < n_full and t * 2 < 6.0 and (16 << (n + 1)) * 2.5 < avail:
        n += 1
        t *= 2
    for _ in range(args.warmup):
        run(n)
    walls = []
    for _ in range(args.steps):
        w, tt, amps = run(n)
        walls.append(w)
    ms = 1e3 * float(np.mean(walls))
    value = amps / (ms / 1e3)
    sample = "QV n=%d depth=%d (%s fits the CPU time budget; full config is n=%d)" % (n, args.depth, name, n_full)
    if args.workload == "qft":
        sample = "QFT n=%d (full config is n=%d)" % (n, n_full)
    print(json.dumps({
        "impl": "reference", "metric": "amplitude_updates_per_s", "value": value, "unit": "amp-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if AMP_BYTES == 16 else "f32", "data": "synthetic",
        "config": {"workload": name, "sample_qubits": n, "depth": args.depth, "fusion_max_qubit": 5,
                   "device": "CPU", "threads": used["threads"] or cores, "host_cores": cores,
                   "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": value, "unit": "amp-updates/s", "cores": used["threads"] or cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "amp-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_leg(args):
    """Reported beside the GPU number (rank 0, N=1): reference Controller on a bounded QV sample."""
    try:
        from oracle import ref_controller
        from qiskit_aer_b200 import circuits
        cores = host_cores()
        if not ref_controller.available():
            raise RuntimeError("oracle/_ref/controller_wrappers.so missing")
        n, t = 22, None
        mk = (lambda m: circuits.qft(m)) if args.workload == "qft" else (lambda m: circuits.quantum_volume(m, args.depth, 1234))
        for _ in range(2):
            t0 = time.perf_counter()
            ref_controller.run_circuit(n, mk(n), shots=SHOTS, seed=1234, threads=cores)
            t = time.perf_counter() - t0
        try:
            avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
        except (ValueError, OSError):
            avail = 32 << 30
        while t * 2 < 8.0 and (16 << (n + 1)) * 2.5 < avail and n < 31:
            n += 1
            t *= 2
        ops = mk(n)
        t0 = time.perf_counter()
        r = ref_controller.run_circuit(n, ops, shots=SHOTS, seed=1234, threads=cores)
        t = time.perf_counter() - t0
        cores = int(r["metadata"].get("parallel_state_update", 0)) or cores  # the threads Aer really used
        return {"value": circuits.amplitudes_written(ops, n) / t, "unit": "amp-updates/s", "cores": cores,
                "kind": "reference",
                "sample": "reference Controller (statevector, device=CPU, fusion_max_qubit=5), %s n=%d, one run, %.2f s wall"
                          % ("QFT" if args.workload == "qft" else "QV depth=%d" % args.depth, n, t)}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": "amp-updates/s", "cores": os.cpu_count(), "kind": "reference",
                "sample": "unavailable: %s" % e}


# ------------------------------------------------------------------------------------------ B200 arm
def b200_arm(args):
    import torch
    import torch.distributed as dist
    import qiskit_aer_b200 as q
    from qiskit_aer_b200 import executor, fusion

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # native libraries (NCCL's version banner) may write to fd 1: park stdout on stderr until the JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    if world > 1 and args.engine == "tile" and args.exchange == "p2p":
        return b200_arm_sharded(args, world, rank, local_rank, dev, saved_stdout)
    n, name, ops, amps_written = workload(args, world)
    if args.precision == "single":
        name += "_f32"
    n_local = n - int(np.log2(world))
    fused = fusion.fuse(ops, max_qubit=args.fusion_max_qubit, max_diag_qubit=args.max_diag_qubit)
    if rank == 0:
        log("workload %s: %d gates -> %d fused passes, n_local=%d (%.1f GiB/GPU)"
            % (name, len(ops), len(fused), n_local, AMP_BYTES * 2.0 ** n_local / 2 ** 30))

    stream = torch.cuda.Stream(device=dev)
    if world > 1 and args.exchange == "p2p":
        # library-owned allocation (exportable over CUDA IPC), kernels ordered on the torch stream NCCL uses
        qv = q.QubitVectorB200(n_local, CDTYPE, device=local_rank)
        qv.set_stream(stream.cuda_stream)
        buf = qv.torch_view()
    else:
        with torch.cuda.stream(stream):
            buf = torch.empty((1 << n_local) * 2, dtype=torch.float64 if AMP_BYTES == 16 else torch.float32, device=dev)
        qv = q.QubitVectorB200(n_local, CDTYPE, device=local_rank, external_ptr=buf.data_ptr(),
                               stream=stream.cuda_stream)
    if world > 1:
        from qiskit_aer_b200 import sharded
        with torch.cuda.stream(stream):
            runner = sharded.ShardedRunner(qv, n, rank, world, stream, buf, exchange=args.exchange)
        plan = runner.plan(fused)
        final_phys = list(runner.phys)
        if rank == 0:
            log("sharded plan: %d global-qubit exchanges" % sum(1 for p in plan if p[0] in ("swap", "mswap")))
    else:
        runner, plan = None, fused

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class_of = lambda op: "diag_layer" if op[0] == "diag_layer" else "%s_k%d" % ("dense" if op[0] == "unitary" else op[0], len(op[1]))  # noqa: E731
    per_class = {}
    launches = [0]
    tile = args.engine == "tile"
    if tile:  # the tile engine consumes the circuit's own 1-/2-qubit gates; passes are formed inside the C ABI
        if runner is not None:
            runner.phys = list(range(n))
            plan = runner.plan(ops)
            final_phys = list(runner.phys)
            if rank == 0:
                log("sharded plan (tile engine): %d exchange steps (%d qubit swaps)" % (
                    sum(1 for p in plan if p[0] in ("swap", "mswap")),
                    sum(1 if p[0] == "swap" else len(p[1]) if p[0] == "mswap" else 0 for p in plan)))
        else:
            plan = ops

    def ev_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run_step(timed):
        if runner is not None:
            runner.initialize()
        else:
            qv.initialize()
        launches[0] += 1
        evs = []
        if tile:
            # time maximal runs of queueable gates as one "tile_pass" sample (launch count = passes used)
            seg = []

            def flush_seg():
                if not seg:
                    return
                st = {}
                if timed:
                    e0, e1 = ev_pair()
                    e0.record(stream)
                executor.apply_ops_queued(qv, seg, st)
                if timed:
                    e1.record(stream)
                    evs.append(("tile_pass", e0, e1, st.get("passes", 0)))
                launches[0] += st.get("launches", 0)
                seg.clear()

            for op in plan:
                if op[0] in ("swap", "mswap"):
                    flush_seg()
                    if timed:
                        e0, e1 = ev_pair()
                        e0.record(stream)
                    launches[0] += runner.apply(op)
                    if timed:
                        e1.record(stream)
                        evs.append((op[0], e0, e1, 1))
                else:
                    seg.append(op)
            flush_seg()
        else:
            for op in plan:
                if timed:
                    e0, e1 = ev_pair()
                    e0.record(stream)
                if runner is not None:
                    launches[0] += runner.apply(op)
                else:
                    executor.apply_op(qv, op)
                    launches[0] += 1
                if timed:
                    e1.record(stream)
                    evs.append((class_of(op) if op[0] in ("unitary", "diagonal", "diag_layer") else op[0], e0, e1, 1))
        if runner is not None:
            runner.phys = list(final_phys)
        return evs

    for _ in range(args.warmup):
        run_step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches[0] = 0
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    all_evs = []
    for _ in range(args.steps):
        all_evs += run_step(True)
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = t_start.elapsed_time(t_end)
    for cls, e0, e1, cnt in all_evs:
        d = per_class.setdefault(cls, [0, 0.0])
        d[0] += cnt
        d[1] += e0.elapsed_time(e1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = amps_written / (ms_per_step / 1e3)
    gpu_launches = launches[0]

    # ---- end to end through the host API: circuit (host) -> fusion -> C ABI -> counts + expvals (host)
    paulis = [([0, 1, n_local - 1], "ZXY"), ([2, 5], "ZZ"), ([3], "X")]

    def e2e_step():
        h2d = sum(executor.op_h2d_bytes(op) for op in ops)
        if tile:
            if runner is not None:
                runner.initialize()
                runner.run_plan(runner.plan(ops))
            else:
                qv.initialize()
                executor.apply_ops_queued(qv, ops)
        else:
            f = fusion.fuse(ops, max_qubit=args.fusion_max_qubit, max_diag_qubit=args.max_diag_qubit)
            h2d = sum(executor.op_h2d_bytes(op) for op in f)
            if runner is not None:
                runner.initialize()
                runner.run_plan(runner.plan(f), queued=False)
            else:
                qv.initialize()
                executor.apply_ops(qv, f)
        rnds = q.rng_uniform(1234, SHOTS)
        if runner is not None:
            samples = runner.sample_measure(rnds)
            ev = [runner.expval_pauli(qs, pl) for qs, pl in paulis]
        else:
            samples = qv.sample_measure(rnds)
            ev = [qv.expval_pauli(qs, pl) for qs, pl in paulis]
        counts = np.unique(samples, return_counts=True)
        return h2d + rnds.nbytes, samples.nbytes + 8 * len(ev), counts, ev

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h2d, d2h, counts, ev = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # ---- the same workload through the reference's own Controller + executors with the B200 vector as
    # its device="GPU" statevector (qiskit-aer_b200/aer/): AerSimulator's C++ stack end to end.
    e2e_aer = None
    if world == 1 and not args.no_aer_e2e:
        try:
            from qiskit_aer_b200 import aer_backend
            if aer_backend.available():
                qv.close()
                del buf
                torch.cuda.empty_cache()
                kw = dict(device="GPU", shots=SHOTS, seed=1234, fusion=False, expvals=paulis, precision=args.precision)
                aer_backend.run_circuit(n, ops, **kw)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    res = aer_backend.run_circuit(n, ops, **kw)
                aer_ms = (time.perf_counter() - t0) * 1e3 / args.steps
                e2e_aer = {"value": amps_written / (aer_ms / 1e3), "unit": "amp-updates/s", "ms_per_step": aer_ms,
                           "time_taken_s": float(res["time_taken"]), "device": res["metadata"].get("device"),
                           "path": "reference Controller::execute -> Statevector::Executor<State<QubitVectorB200<double>>> "
                                   "(fusion_enable=false: the adapter's gate queue feeds the tile engine)"}
        except Exception as e:  # reported, never fatal
            e2e_aer = {"value": None, "error": str(e)[:300]}

    peak64 = fp64_peak(local_rank, 1000.0) if (AMP_BYTES == 16 and tile) else None
    extras = {}
    if world == 1 and rank == 0 and not args.no_extras and args.workload == "qv":
        try:
            qv.close()
        except Exception:
            pass
        buf = None
        torch.cuda.empty_cache()
        from qiskit_aer_b200 import capi
        capi.lib().b200sv_trim()   # the child processes of the head-to-head need the memory the library keeps for reuse
        for key, fn in (("qft30", lambda: extra_qft30(local_rank)), ("noisy20_10k", extra_noisy20),
                        ("qv33_single_precision", extra_qv33_single), ("vs_reference_gpu", extra_vs_reference_gpu)):
            t0 = time.perf_counter()
            try:
                extras[key] = fn()
            except Exception as e:  # reported, never fatal
                extras[key] = {"error": str(e)[:300]}
            extras[key]["leg_seconds"] = time.perf_counter() - t0
            torch.cuda.empty_cache()
            capi.lib().b200sv_trim()

    if rank == 0:
        peak, peak_src = measured_peak()
        dom = max((c for c in per_class if c.startswith(("dense", "diag", "tile"))), key=lambda c: per_class[c][1])
        cnt, tot = per_class[dom]
        avg_ms = tot / cnt
        bytes_per_launch = 2 * AMP_BYTES * 2.0 ** n_local
        achieved = bytes_per_launch / (avg_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tp):
            ent = json.load(open(tp)).get(dom)
            if ent:  # per-amplitude DRAM bytes from the committed ncu --set full capture, scaled to this launch
                traffic = ent["bytes_per_amp"] * 2.0 ** n_local
        fp64 = None
        if dom == "tile_pass" and AMP_BYTES == 16:
            # the tile pass carries several gates per HBM pass and is FP64-pipe bound: report that roofline too,
            # against the DFMA peak measured NOW on this GPU (sustained = under the same power cap as the passes)
            flops = sum(8.0 * 2 ** len(op[1] if op[0] == "unitary" else op[2]) for op in ops) * 2.0 ** n_local
            tf = flops * args.steps / (tot / 1e3) / 1e12
            fp64 = {"achieved_tflops": tf, "peak_tflops": peak64["sustained_tflops"], "frac": tf / peak64["sustained_tflops"],
                    "peak_burst_tflops": peak64["burst_tflops"],
                    "peak_source": "b200sv_measure_fp64_peak, measured in this run on this GPU (DFMA, constant-bank operands)",
                    "gates_per_pass": len(ops) * args.steps / max(cnt, 1)}
        out = {
            "metric": "amplitude_updates_per_s", "value": value, "unit": "amp-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if AMP_BYTES == 16 else "f32", "data": "synthetic",
            "config": {"workload": name, "qubits": n, "qubits_per_gpu": n_local, "depth": args.depth,
                       "circuit_gates": len(ops), "engine": args.engine,
                       "hbm_passes": (per_class["tile_pass"][0] // args.steps) if tile else len(fused),
                       "fusion": ("tile-blocked gate queue: 2^12-amplitude shared-memory tiles, gates applied from registers"
                                  if tile else "dense blocks, max_qubit=%d" % args.fusion_max_qubit),
                       "shots": SHOTS,
                       "l2": "state (%.0f GiB per GPU) is larger than L2; no flush needed" % (AMP_BYTES * 2.0 ** n_local / 2 ** 30),
                       "sharding": "top %d qubits select the GPU" % int(np.log2(world)),
                       "exchange": (args.exchange if world > 1 else None)},
            "wall_time_s": ms_per_step / 1e3,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": dom, "launches": cnt, "avg_ms": avg_ms,
                         "bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                         "frac_of_nominal_8000": achieved / 8000.0, "fp64": fp64,
                         # the north star's "per-gate GB/s": what one streaming pass PER GATE would have to sustain to
                         # finish the circuit in the same time (several gates share one HBM pass in the tile engine)
                         "per_gate_equivalent": {
                             "GBps": AMP_BYTES * 2.0 * amps_written / world / (ms_per_step / 1e3) / 1e9,
                             "x_hbm_peak": AMP_BYTES * 2.0 * amps_written / world / (ms_per_step / 1e3) / 1e9 / peak}},
            "per_kernel": {c: {"launches": v[0], "avg_ms": v[1] / v[0],
                               "GBps": (bytes_per_launch / (v[1] / v[0] / 1e3) / 1e9) if c.startswith(("dense", "diagonal", "tile")) else None}
                           for c, v in sorted(per_class.items())},
            "e2e": {"value": amps_written / (e2e_ms / 1e3), "unit": "amp-updates/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
        }
        if e2e_aer is not None:
            out["e2e_aer_controller"] = e2e_aer
        if extras:
            out["other_configs"] = extras
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg(args)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------ the other BASELINE configs
def extra_qft30(local_rank):
    """BASELINE config 2: QFT-30, double.  (a) the engine's own front end (commutation-aware fusion, wide diagonal
    passes, gate queue), (b) the reference Controller with fusion_max_qubit=5 on the B200 vector."""
    import torch
    import qiskit_aer_b200 as q
    from qiskit_aer_b200 import aer_backend, circuits, executor, fusion
    n = 30
    ops = circuits.qft(n)
    amps = circuits.amplitudes_written(ops, n)
    # commutation-aware fusion: dense blocks <= 4 qubits, diagonal tables <= 16 qubits, wider sets of commuting
    # controlled phases as ONE streaming pass each (b200sv_apply_diagonal_layer)
    fused = fusion.fuse(ops, max_qubit=4, max_diag_qubit=40, max_table_qubit=16)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        buf = torch.empty((1 << n) * 2, dtype=torch.float64, device="cuda:%d" % local_rank)
    qv = q.QubitVectorB200(n, np.complex128, device=local_rank, external_ptr=buf.data_ptr(), stream=stream.cuda_stream)
    def cls(op):
        if op[0] == "diag_layer":
            return "diag_layer"
        return "%s_k%d" % ("dense" if op[0] == "unitary" else op[0], len(op[1])) if op[0] in ("unitary", "diagonal") else op[1]
    per = {}
    walls = []
    for rep in range(4):
        qv.initialize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs = []
        e0.record(stream)
        for op in fused:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            executor.apply_op(qv, op)
            b.record(stream)
            evs.append((cls(op), a, b))
        e1.record(stream)
        torch.cuda.synchronize()
        if rep:
            walls.append(e0.elapsed_time(e1))
            for c, a, b in evs:
                d = per.setdefault(c, [0, 0.0])
                d[0] += 1
                d[1] += a.elapsed_time(b)
    ev_own = qv.expval_pauli([0, 1, n - 1], "ZXY")
    # the same circuit handed over gate by gate as ONE queue (b200sv_apply_gate_sequence: the library regroups it into
    # dense blocks -> tile passes and wide diagonal layers; what the adapter's queue flush does under the Controller)
    queue_ops = [("unitary", list(o[2]), fusion.gate_matrix(o[1], o[3])) for o in ops]
    qwalls, qstats = [], {}
    for rep in range(4):
        qv.initialize()
        qstats = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        executor.apply_ops_queued(qv, queue_ops, stats=qstats)
        e1.record(stream)
        torch.cuda.synchronize()
        if rep:
            qwalls.append(e0.elapsed_time(e1))
    ev_queue = qv.expval_pauli([0, 1, n - 1], "ZXY")
    qv.close()
    del buf
    torch.cuda.empty_cache()
    dom = max(per, key=lambda c: per[c][1])
    avg = per[dom][1] / per[dom][0]
    peak, src = measured_peak()
    bytes_per_launch = 2 * 16 * 2.0 ** n
    best = min(float(np.mean(walls)), float(np.mean(qwalls)))
    out = {"workload": "qft30_fused", "qubits": n, "circuit_gates": len(ops), "hbm_passes": len(fused),
           "ms_per_circuit": best, "amp_updates_per_s": amps / (best / 1e3),
           "ms_op_by_op_front_end": float(np.mean(walls)),
           "gate_queue": {"ms_per_circuit": float(np.mean(qwalls)), "hbm_passes": int(qstats.get("passes", 0)),
                          "expval_agrees": bool(abs(ev_queue - ev_own) < 1e-10)},
           "roofline": {"bound": "hbm", "kernel": dom, "launches": per[dom][0], "avg_ms": avg,
                        "achieved": bytes_per_launch / (avg / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": bytes_per_launch / (avg / 1e3) / 1e9 / peak, "peak_source": src,
                        "share_of_circuit": per[dom][1] / sum(v[1] for v in per.values())}}
    if aer_backend.available():
        kw = dict(device="GPU", shots=SHOTS, seed=1234, fusion=True, fusion_max_qubit=5, fusion_threshold=14,
                  expvals=[([0, 1, n - 1], "ZXY")])
        aer_backend.run_circuit(n, ops, **kw)
        t0 = time.perf_counter()
        r = aer_backend.run_circuit(n, ops, **kw)
        dt = time.perf_counter() - t0
        kw0 = dict(kw, fusion=False)
        aer_backend.run_circuit(n, ops, **kw0)
        t0 = time.perf_counter()
        r0 = aer_backend.run_circuit(n, ops, **kw0)
        dt0 = time.perf_counter() - t0
        out["aer_controller_fusion5"] = {"wall_s": dt, "time_taken_s": float(r["time_taken"]),
                                         "expval_agrees": bool(abs(float(r["data"]["ev0"]) - ev_own) < 1e-10)}
        out["aer_controller_gate_queue"] = {"wall_s": dt0, "time_taken_s": float(r0["time_taken"]),
                                            "expval_agrees": bool(abs(float(r0["data"]["ev0"]) - ev_own) < 1e-10)}
    return out


def extra_noisy20():
    """BASELINE config 5: noisy 20-qubit circuit, depolarizing noise, 10 k shots through the reference's
    BatchShotsExecutor (batched_shots_gpu) on the B200 container; Pauli expectation values checked against the
    reference CPU path (per-shot, bounded shots) within 5 sigma."""
    from qiskit_aer_b200 import aer_backend, circuits, noise
    if not aer_backend.available():
        return {"unavailable": "Aer integration module not built"}
    n, depth, shots, cpu_shots = 20, 20, 10000, 96
    ops = circuits.random_noisy_circuit(n, depth, seed=1)
    rng = np.random.default_rng(0)
    obs = []
    for _ in range(10):
        k = int(rng.integers(1, 5))
        qs = [int(x) for x in rng.choice(n, size=k, replace=False)]
        obs.append((qs, "".join("XYZ"[int(c)] for c in rng.integers(0, 3, size=k))))
    nm = noise.noise_model_dict(1e-3, 1e-2)
    kw = dict(seed=3, fusion=False, noise_model=nm, expvals=obs)
    bkw = dict(kw, batched_shots_gpu=True, batched_shots_gpu_max_qubits=max(n, 16))
    aer_backend.run_circuit(n, ops, device="GPU", shots=64, **bkw)
    t0 = time.perf_counter()
    r = aer_backend.run_circuit(n, ops, device="GPU", shots=shots, **bkw)
    dt = time.perf_counter() - t0
    ev_gpu = np.array([float(r["data"]["ev%d" % i]) for i in range(len(obs))])
    t0 = time.perf_counter()
    # per-shot values on the CPU side ("list"): their spread gives the standard error of both shot means
    c = aer_backend.run_circuit(n, ops, device="CPU", shots=cpu_shots, threads=host_cores(), expval_subtype="list", **kw)
    dtc = time.perf_counter() - t0
    per_shot = np.array([[float(v) for v in c["data"]["ev%d" % i]] for i in range(len(obs))])  # [obs][shot]
    ev_cpu = per_shot.mean(axis=1)
    std = per_shot.std(axis=1, ddof=1)
    sigma = np.maximum(std * np.sqrt(1.0 / shots + 1.0 / cpu_shots), 1e-12)
    dev_sigma = float(np.max(np.abs(ev_gpu - ev_cpu) / sigma))
    return {"workload": "noisy_random_circuit", "qubits": n, "depth": depth, "gates": len(ops), "shots": shots,
            "p1": 1e-3, "p2": 1e-2, "observables": len(obs), "seconds": dt, "shots_per_s": shots / dt,
            "batched_shots_optimization": bool(r["metadata"].get("batched_shots_optimization")),
            "reference_cpu": {"shots": cpu_shots, "seconds": dtc, "shots_per_s": cpu_shots / dtc,
                              "threads": max(int(c["metadata"].get("parallel_state_update", 1)),
                                             int(c["metadata"].get("parallel_shots", 1)))},
            "expval_gpu": [float(x) for x in ev_gpu], "expval_cpu": [float(x) for x in ev_cpu],
            "expval_sigma": [float(x) for x in sigma],
            "expval_max_deviation_sigma": dev_sigma, "expval_within_5_sigma": bool(dev_sigma < 5.0)}


def extra_qv33_single():
    """The headline circuit in single precision (complex64, 64 GiB): the float tile passes, whose rounds run on packed
    FP32 FMAs (FFMA2).  The same bench in a child process (`--precision single`), two timed steps."""
    cmd = [sys.executable, os.path.abspath(__file__), "--precision", "single", "--steps", "2", "--warmup", "3",
           "--no-cpu-baseline", "--no-aer-e2e", "--no-extras"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not line:
        return {"error": (r.stderr or r.stdout)[-300:]}
    d = json.loads(line[-1])
    rf = d.get("roofline", {})
    return {"workload": d["config"]["workload"], "dtype": d["dtype"], "ms_per_step": d["ms_per_step"],
            "amp_updates_per_s": d["value"], "hbm_passes": d["config"].get("hbm_passes"),
            "tile_pass_avg_ms": rf.get("avg_ms"), "hbm_frac": rf.get("frac"), "e2e_ms_per_step": d.get("e2e", {}).get("ms_per_step"),
            "clocks": d.get("clocks")}


def extra_vs_reference_gpu():
    """The reference's OWN GPU path (QubitVectorThrust, compiled unmodified for sm_100: oracle/_ref/gpu) against the
    same Controller on the B200 engine, identical circuit / seed / options, each side in its own process."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_vs_reference_gpu as h2h
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gpu", "controller_wrappers.so")):
        return {"unavailable": "oracle/_ref/gpu/controller_wrappers.so not built (make -C oracle ref-gpu)"}
    n = 31

    def side(name, fusion):
        smp = ClockSampler(0)
        smp.start()
        r = h2h.run(name, n, "qv", fusion, 5)
        r["clocks"] = smp.stop()
        return r

    same = side("b200_engine", True)
    ref = side("reference_gpu", True)
    best = side("b200_engine", False)
    out = {"workload": "qv31_depth10", "reference_thrust_sm100_fusion5": ref, "b200_engine_same_options": same,
           "b200_engine_gate_queue": best}
    try:
        out["speedup_same_options"] = ref["wall_s"] / same["wall_s"]
        out["speedup_best"] = ref["wall_s"] / best["wall_s"]
        out["expval_agree"] = bool(abs(ref["ev"] - same["ev"]) < 1e-9 and abs(ref["ev"] - best["ev"]) < 1e-9)
    except KeyError:
        pass
    return out


# ------------------------------------------------------------------------------------------ B200 arm, N > 1
def fp64_peak(device, ms=300.0):
    """DFMA issue peak of this device, measured now (b200sv_measure_fp64_peak)."""
    import ctypes as C
    from qiskit_aer_b200 import capi
    b, s = C.c_double(0), C.c_double(0)
    capi.check(capi.lib().b200sv_measure_fp64_peak(int(device), float(ms), C.byref(b), C.byref(s)))
    return {"burst_tflops": b.value, "sustained_tflops": s.value}


def b200_arm_sharded(args, world, rank, local_rank, dev, saved_stdout):
    """One process per GPU; the register is driven by the C++ sharded executor (csrc/sharded.cu): un-fused circuit
    gates -> epoch plan -> tile passes on every shard -> staged, slab-pipelined exchanges over NVLink.  Python only
    encodes the circuit, moves IPC handles / scalars with torch.distributed and reads CUDA events."""
    import torch
    import torch.distributed as dist
    import qiskit_aer_b200 as q
    from qiskit_aer_b200 import circuits, executor, sharded

    n, name, ops, amps_written = workload(args, world)
    if args.precision == "single":
        name += "_f32"
    g = int(np.log2(world))
    n_local = n - g

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # ---- parity check on the SAME code path at a size one GPU also holds unsharded: sharded run == this rank's own
    # unsharded run of the same circuit (chunked == unchunked, test/terra/backends/aer_simulator/test_chunk.py:31-168)
    n_chk = 20 + g
    chk_ops = circuits.quantum_volume(n_chk, args.depth, seed=4321)
    saved_env = {k: os.environ.get(k) for k in ("B200SV_SHARD_MIN_RUN_BITS", "B200SV_SHARD_SLAB_BITS")}
    os.environ["B200SV_SHARD_MIN_RUN_BITS"] = "6"
    os.environ["B200SV_SHARD_SLAB_BITS"] = "3"
    stc = sharded.ShardedState(n_chk, world=world, rank=rank, device=local_rank, dist=dist, dtype=CDTYPE,
                               staging_bytes=(16 << (n_chk - g)) // 4)
    for k, v in saved_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    stc.initialize()
    stc.apply_ops(chk_ops)
    chk_stats = stc.stats()
    chk_norm = stc.norm()
    stc.restore_order()
    mine = stc.shard_vector(rank)
    one = q.QubitVectorB200(n_chk, CDTYPE, device=local_rank)
    one.initialize()
    executor.apply_ops_queued(one, chk_ops)
    want = one.vector()[rank << (n_chk - g):(rank + 1) << (n_chk - g)]
    err = float(np.max(np.abs(mine - want)))
    one.close()
    stc.close()
    from qiskit_aer_b200 import capi
    capi.lib().b200sv_trim()
    tol = 1e-12 if AMP_BYTES == 16 else 5e-6
    t = torch.tensor([err, abs(chk_norm - 1.0)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    parity = {"path": "b200sv_sharded_apply_ops (tile passes on shards + staged exchange), same as the timed loop",
              "against": "unsharded run of the same circuit on one GPU (b200sv_apply_gate_sequence)",
              "qubits": n_chk, "gates": len(chk_ops), "exchanges": chk_stats["exchanges"], "staged": chk_stats["staged"],
              "overlapped_passes": chk_stats["overlapped_passes"], "max_abs_amp_err": float(t[0].item()),
              "norm_err": float(t[1].item()), "tol": tol, "ok": bool(t[0].item() < tol and t[1].item() < 1e-10)}
    if not parity["ok"]:
        raise SystemExit("bench.py: sharded parity check FAILED: %s" % json.dumps(parity))

    # ---- the timed workload
    st = sharded.ShardedState(n, world=world, rank=rank, device=local_rank, dist=dist, dtype=CDTYPE)
    enc = st.encode(ops)
    stream = torch.cuda.ExternalStream(st.compute_stream(), device=dev)
    if rank == 0:
        log("workload %s: %d gates, n_local=%d (%.1f GiB/GPU), plan: %s" % (
            name, len(ops), n_local, AMP_BYTES * 2.0 ** n_local / 2 ** 30,
            sharded.ShardedState.plan_only(n, world, ops, 34 << 30, dtype=CDTYPE)))

    def run_step():
        st.initialize()
        st.apply_ops(encoded=enc)

    for _ in range(args.warmup):
        run_step()
    st.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    st.profile(True)
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for _ in range(args.steps):
        run_step()
    t_end.record(stream)
    st.synchronize()
    barrier()
    clocks = sampler.stop()
    prof = st.profile_read()
    st.profile(False)
    stats = st.stats()
    t = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = amps_written / (ms_per_step / 1e3)
    norm_after = st.norm()
    if abs(norm_after - 1.0) > (1e-10 if AMP_BYTES == 16 else 1e-4):
        raise SystemExit("bench.py: global norm after the timed loop is %r" % norm_after)

    # ---- end to end through the host API: circuit in host memory -> encode -> C ABI -> counts + expvals on the host
    paulis = [([0, 1, n - 1], "ZXY"), ([2, 5], "ZZ"), ([3], "X")]

    def e2e_step():
        e = st.encode(ops)
        st.initialize()
        st.apply_ops(encoded=e)
        ev = [st.expval_pauli(qs, pl) for qs, pl in paulis]      # before the order is restored: no data moves
        rnds = q.rng_uniform(1234, SHOTS)
        samples = st.sample_measure(rnds)
        return int(e[4].nbytes + e[2].nbytes + rnds.nbytes), int(samples.nbytes + 8 * len(ev)), samples, ev

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        h2d, d2h, samples, ev = e2e_step()
    barrier()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    peak64 = fp64_peak(local_rank) if AMP_BYTES == 16 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        tp = prof["tile_pass"]
        avg_ms = tp["ms"] / max(tp["count"], 1)
        bytes_per_launch = 2 * AMP_BYTES * 2.0 ** n_local
        achieved = bytes_per_launch / (avg_ms / 1e3) / 1e9 if avg_ms else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tpath):
            ent = json.load(open(tpath)).get("tile_pass")
            if ent:
                traffic = ent["bytes_per_amp"] * 2.0 ** n_local
        flops = sum(8.0 * 2 ** len(op[1]) for op in ops) * 2.0 ** n_local  # QV: dense 2-qubit gates
        passes_per_step = stats["passes"]
        tile_total_ms = avg_ms * passes_per_step * args.steps
        fp64 = None
        if peak64 is not None and tile_total_ms:
            tf = flops * args.steps / (tile_total_ms / 1e3) / 1e12
            fp64 = {"achieved_tflops": tf, "peak_tflops": peak64["sustained_tflops"], "frac": tf / peak64["sustained_tflops"],
                    "peak_burst_tflops": peak64["burst_tflops"],
                    "peak_source": "b200sv_measure_fp64_peak, measured in this run on this GPU (DFMA, constant-bank operands)",
                    "gates_per_pass": len(ops) / max(passes_per_step, 1)}
        # exchange: what crosses NVLink, how long the copy engines take, how much of it the passes hide
        nx = max(prof["exchange_region"]["count"] + prof["exchange_inplace"]["count"], 1)
        region_ms = prof["exchange_region"]["ms"] / max(prof["exchange_region"]["count"], 1)
        push_ms = prof["push"]["ms"] / nx
        slab_pass_ms = prof["slab_pass"]["ms"] / nx
        unstage_ms = prof["unstage"]["ms"] / nx
        bytes_x = stats["bytes_sent_per_shard"] / max(stats["exchanges"], 1)
        exposed_ms = max(region_ms - slab_pass_ms, 0.0) if prof["exchange_region"]["count"] else prof["exchange_inplace"]["ms"] / nx
        exchange = {"per_step": stats["exchanges"], "staged": stats["staged"], "inplace": stats["inplace"],
                    "qubit_swaps_per_step": None, "bytes_per_direction_per_gpu": bytes_x,
                    "push_ms": push_ms, "nvlink_GBps_per_direction": bytes_x / (push_ms / 1e3) / 1e9 if push_ms else None,
                    "nvlink_peak_GBps": 770.0, "nvlink_peak_source": "measured peer copy (B200_PROFILING.md), 900 nominal",
                    "region_ms": region_ms, "slab_pass_ms_inside_region": slab_pass_ms, "unstage_ms": unstage_ms,
                    "exposed_ms": exposed_ms,
                    "overlap_fraction": (1.0 - min(1.0, max(exposed_ms - unstage_ms, 0.0) / push_ms)) if push_ms else 0.0,
                    "overlapped_passes_per_step": stats["overlapped_passes"], "dma_copies_per_step": stats["copies"]}
        out = {
            "metric": "amplitude_updates_per_s", "value": value, "unit": "amp-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if AMP_BYTES == 16 else "f32", "data": "synthetic",
            "config": {"workload": name, "qubits": n, "qubits_per_gpu": n_local, "depth": args.depth,
                       "circuit_gates": len(ops), "engine": "tile", "hbm_passes": passes_per_step,
                       "fusion": "tile-blocked gate queue: 2^12-amplitude shared-memory tiles, gates applied from registers",
                       "shots": SHOTS,
                       "l2": "state (%.0f GiB per GPU) is larger than L2; no flush needed" % (AMP_BYTES * 2.0 ** n_local / 2 ** 30),
                       "sharding": "top %d qubits select the GPU" % g,
                       "host": "C++ sharded executor (b200sv_sharded_apply_ops): epoch plan, tile passes, staged "
                               "slab-pipelined exchange (copy-engine pushes over NVLink, flag-word ordering)",
                       "exchange": "staged-dma"},
            "wall_time_s": ms_per_step / 1e3,
            "parity_check": parity, "norm_after_timed_loop": norm_after,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "traffic": traffic, "kernel": "tile_pass",
                         "launches": tp["count"], "avg_ms": avg_ms, "bytes_per_launch": bytes_per_launch,
                         "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0 if achieved else None,
                         "fp64": fp64,
                         "per_gate_equivalent": {
                             "GBps": AMP_BYTES * 2.0 * amps_written / world / (ms_per_step / 1e3) / 1e9,
                             "x_hbm_peak": AMP_BYTES * 2.0 * amps_written / world / (ms_per_step / 1e3) / 1e9 / peak}},
            "per_kernel": {k: {"launches": v["count"], "avg_ms": v["ms"] / v["count"]} for k, v in prof.items() if v["count"]},
            "exchange": exchange,
            "e2e": {"value": amps_written / (e2e_ms / 1e3), "unit": "amp-updates/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
            "gpu_launches": int(stats["launches"] * args.steps),
            "clocks": clocks,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    st.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qv", choices=["qv", "qft"])
    ap.add_argument("--qubits", type=int, default=0, help="override the total qubit count (default 33 + log2 N)")
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--fusion-max-qubit", type=int, default=4)
    ap.add_argument("--max-diag-qubit", type=int, default=40,
                    help="widest fused diagonal block (2^k table up to 16 qubits, one-pass diagonal layer above)")
    ap.add_argument("--engine", default="tile", choices=["tile", "dense"],
                    help="tile: multi-gate shared-memory passes (default); dense: one fused dense block per pass")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="global-qubit exchange: in-place NVLink peer swap kernel over CUDA IPC, or ncclSend/Recv slices")
    ap.add_argument("--precision", default="double", choices=["double", "single"],
                    help="amplitude type; BASELINE's metric is quoted in double precision (the default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aer-e2e", action="store_true", help="skip the run through the reference Controller")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the other BASELINE configs (QFT-30, noisy 20 q / 10 k shots, reference GPU head-to-head)")
    args = ap.parse_args()
    global AMP_BYTES, CDTYPE
    AMP_BYTES = 16 if args.precision == "double" else 8
    CDTYPE = np.complex128 if args.precision == "double" else np.complex64
    if args.workload == "qft":
        args.engine = "dense"  # QFT is mostly controlled phases: commutation-aware fusion + wide diagonal passes
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()

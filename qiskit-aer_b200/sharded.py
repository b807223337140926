"""Sharded statevector: one process per GPU, state split by the top log2(G) qubits.

Replaces the reference's chunk distribution + global-qubit exchange
(/root/reference/src/simulators/parallel_state_executor.hpp: apply_ops_chunks :772,
apply_chunk_swap :1134-1336, MPI_Isend/Irecv :1317-1327) and the cache-blocking transpile pass that
schedules it (src/transpile/cacheblocking.hpp:182-203,385-720), redesigned for NVSwitch:

* one chunk per GPU (chunk_bits = n - log2 G): 3 global qubits for QV-36 on 8 GPUs instead of the
  reference's >= 6 (it needs an exchange buffer chunk, chunk_manager.hpp:281);
* NO swap-back: a logical->physical qubit map is carried on the host (all ranks compute the same
  plan deterministically, no communication), so a global qubit that is swapped in stays local until
  evicted; the victim is the local qubit whose next use is farthest away (Belady), restricted to high
  physical positions so every transfer slice is one contiguous run (zero-copy send);
* diagonal gates and controls on global qubits never move data (resolved from the chunk index,
  thrust_kernels.hpp:1190,1342,2004 `base_index_` rule -- done inside the C ABI);
* the exchange itself is pairwise ncclSend/ncclRecv (torch.distributed P2P ops) in slices, received
  into a double-buffered staging area and copied into place while the next slice is in flight.

Sampling / expectation values follow Statevector::Executor::{sample_measure :1149, expval_pauli :551}
(statevector_executor.hpp): per-rank partial results, reduced over ranks; sampled physical indices are
mapped back to logical bit order on the host.
"""
import numpy as np
import torch
import torch.distributed as dist


def _insert_zero(v, pos):
    low = v & ((1 << pos) - 1)
    return ((v >> pos) << (pos + 1)) | low


class ShardedRunner:
    def __init__(self, qv, n, rank, world, stream, buf, slice_amps=1 << 26, min_run_bits=20, exchange="nccl"):
        """qv: chunk backend (QubitVectorB200 or a CPU stand-in with the same methods);
        buf: torch tensor viewing the chunk's amplitudes as float64 pairs (exchange source/target)."""
        self.qv, self.n, self.rank, self.world = qv, int(n), int(rank), int(world)
        self.gbits = int(np.log2(world))
        if (1 << self.gbits) != world:
            raise ValueError("number of GPUs must be a power of two")
        self.nl = self.n - self.gbits
        self.stream, self.buf = stream, buf
        self.amps = buf.view(-1, 2)  # [2^nl, 2]
        self.slice_amps = int(min(slice_amps, 1 << max(self.nl - 1, 0)))
        self.min_run_bits = min(min_run_bits, max(self.nl - 1, 0))
        self.phys = list(range(self.n))  # logical qubit -> physical position
        self._tmp = None
        qv.chunk_setup(self.n, self.rank)
        self.bytes_exchanged = 0
        self.exchange = exchange
        self.multi_swap = True   # fold the swaps of one epoch boundary into one all-to-all pass
        self._peer_ptr = {}
        if exchange == "p2p":
            # CUDA-IPC: map every exchange partner's chunk into this process (NVLink peer access)
            handles = [None] * world
            dist.all_gather_object(handles, qv.ipc_export())
            self._rank_ptr = {}
            for r in range(world):
                if r != self.rank:
                    self._rank_ptr[r] = qv.ipc_open(handles[r])
            for gb in range(self.gbits):
                self._peer_ptr[gb] = self._rank_ptr[self.rank ^ (1 << gb)]
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.amps.device)

    # ---------------------------------------------------------------- planning (host, deterministic)
    def plan(self, ops, phys=None):
        """Rewrite logical ops to physical ones, inserting ("swap", local_pos, global_bit) / ("mswap", ...) steps.
        The scheduling itself is b200sv_plan_epochs (csrc/planner.cu); `_plan_py` below is the same algorithm in
        Python, kept as the cross-check of tests/test_host_logic.py."""
        import ctypes as C
        from . import capi
        phys = list(self.phys if phys is None else phys)
        off, qs, need = [0], [], []
        for op in ops:
            q = list(op[1]) if op[0] in ("unitary", "diagonal") else list(op[2])
            if op[0] == "diagonal" or (op[0] == "gate" and op[1] == "cp"):
                loc = []
            else:
                loc = list(op[1]) if op[0] == "unitary" else self._gate_targets(op)
            qs += q
            need += [1 if x in loc else 0 for x in q]
            off.append(len(qs))
        n_ops = len(ops)
        a_off = np.asarray(off, dtype=np.int32)
        a_qs = np.asarray(qs if qs else [0], dtype=np.int32)
        a_need = np.asarray(need if need else [0], dtype=np.uint8)
        a_phys = np.asarray(phys, dtype=np.int32)
        cap = len(qs) + 13 * (n_ops + 1)
        buf = np.zeros(cap, dtype=np.int64)
        plen = C.c_int64(0)
        capi.check(capi.lib().b200sv_plan_epochs(
            self.n, self.nl, n_ops, a_off.ctypes.data_as(C.POINTER(C.c_int)), a_qs.ctypes.data_as(C.POINTER(C.c_int)),
            a_need.ctypes.data_as(C.POINTER(C.c_uint8)), int(self.min_run_bits), 1 if self.multi_swap else 0,
            a_phys.ctypes.data_as(C.POINTER(C.c_int)), buf.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(plen)))
        out, i = [], 0
        rec = buf[:plen.value].tolist()
        while i < len(rec):
            if rec[i] == 0:
                op, k = ops[rec[i + 1]], rec[i + 2]
                pq = rec[i + 3:i + 3 + k]
                if op[0] == "unitary":
                    out.append(("unitary", pq, op[2]))
                elif op[0] == "diagonal":
                    out.append(("diagonal", pq, op[2]))
                else:
                    out.append(("gate", op[1], pq, op[3]))
                i += 3 + k
            elif rec[i] == 1:
                out.append(("swap", rec[i + 1], rec[i + 2]))
                i += 3
            else:
                k = rec[i + 1]
                out.append(("mswap", rec[i + 2:i + 2 + k], rec[i + 2 + k:i + 2 + 2 * k]))
                i += 2 + 2 * k
        self.phys = [int(x) for x in a_phys]
        return out

    def _plan_py(self, ops, phys=None):
        """Reference implementation of plan() (same algorithm, pure Python).

        Epoch scheduling: run EVERY gate that is executable under the current qubit map (respecting
        dependencies through shared qubits), then bring in the global qubit(s) the blocked gates wait
        for, evicting the local qubits whose next use is farthest away (Belady), and repeat.  Long
        epochs matter twice: fewer exchanges, and the tile engine packs a whole epoch's gates into
        few HBM passes.  Diagonal gates and controls never block on a global qubit (they are resolved
        from the chunk index).  Updates self.phys to the mapping valid after the plan has run."""
        phys = list(self.phys if phys is None else phys)
        nl = self.nl

        def qubits_of(op):
            return list(op[1]) if op[0] in ("unitary", "diagonal") else list(op[2])

        def need_local(op):
            if op[0] == "diagonal" or (op[0] == "gate" and op[1] == "cp"):
                return []
            return list(op[1]) if op[0] == "unitary" else self._gate_targets(op)

        def emit(op):
            qs = [phys[q] for q in qubits_of(op)]
            if op[0] == "unitary":
                return ("unitary", qs, op[2])
            if op[0] == "diagonal":
                return ("diagonal", qs, op[2])
            return ("gate", op[1], qs, op[3])

        out = []
        remaining = list(range(len(ops)))
        while remaining:
            blocked, rest, wanted, frontier = set(), [], [], set()
            for i in remaining:
                qs = qubits_of(ops[i])
                if blocked.intersection(qs):
                    blocked.update(qs)
                    rest.append(i)
                    continue
                glob = [q for q in need_local(ops[i]) if phys[q] >= nl]
                if glob:
                    blocked.update(qs)
                    rest.append(i)
                    frontier.update(qs)  # never evict a partner of the gate we are swapping for
                    for q in glob:
                        if q not in wanted:
                            wanted.append(q)
                    continue
                out.append(emit(ops[i]))
            remaining = rest
            if not remaining:
                break
            # next use (position in `remaining`) of every logical qubit that must be local there
            uses = {}
            for pos, i in enumerate(remaining):
                for q in need_local(ops[i]):
                    uses.setdefault(q, []).append(pos)
            busy = set(wanted) | frontier
            pairs = []
            for q in wanted[:self.gbits]:
                victim = self._pick_victim(phys, busy, uses, -1)
                busy.add(victim)
                lpos, gpos = phys[victim], phys[q]
                pairs.append((lpos, gpos - nl))
                phys[victim], phys[q] = gpos, lpos
            if len(pairs) > 1 and self.multi_swap:
                out.append(("mswap", [a for a, _ in pairs], [b for _, b in pairs]))
            else:
                out.extend(("swap", a, b) for a, b in pairs)
        self.phys = phys
        return out

    @staticmethod
    def _gate_targets(op):
        name, qs = op[1], op[2]
        if name == "swap":
            return qs[-2:]
        return qs[-1:]  # controls may stay global

    def _pick_victim(self, phys, busy, uses, now):
        nl = self.nl
        inv = {p: q for q, p in enumerate(phys)}
        cands = [p for p in range(nl - 1, -1, -1) if inv[p] not in busy]
        high = [p for p in cands if p >= self.min_run_bits]
        cands = high or cands
        best, best_next = None, -1
        for p in cands:
            nxt = next((u for u in uses.get(inv[p], []) if u > now), 1 << 60)
            if nxt > best_next:
                best, best_next = p, nxt
        if best is None:
            raise RuntimeError("no local qubit available to evict")
        return inv[best]

    # ---------------------------------------------------------------- execution
    def initialize(self):
        """|0...0> of the whole register: chunk 0 holds the 1, every other chunk is zero
        (Statevector::Executor::initialize_qreg, statevector_executor.hpp:437-470)."""
        self.phys = list(range(self.n))
        if self.rank == 0:
            self.qv.initialize()
        else:
            self.qv.zero()

    def _ctx(self):
        import contextlib
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def apply(self, op):
        """Returns the number of kernel launches issued."""
        from .executor import apply_op
        if op[0] == "swap":
            with self._ctx():
                return self.swap_global(op[1], op[2])
        if op[0] == "mswap":
            with self._ctx():
                return self.multi_swap_global(op[1], op[2])
        apply_op(self.qv, op)
        return 1

    def multi_swap_global(self, lposs, gbits):
        """k local qubits <-> k global bits in ONE in-place pass over the peer mappings (all-to-all among the
        2^k ranks that differ in those bits); falls back to k pairwise exchanges without peer mappings."""
        if self.exchange != "p2p":
            return sum(self.swap_global(a, b) for a, b in zip(lposs, gbits))
        k = len(lposs)
        my_g = 0
        for i, gb in enumerate(gbits):
            my_g |= ((self.rank >> gb) & 1) << i
        peers = []
        for v in range(1 << k):
            r = self.rank
            for i, gb in enumerate(gbits):
                r = (r & ~(1 << gb)) | (((v >> i) & 1) << gb)
            peers.append(self._rank_ptr.get(r, 0))
        dist.all_reduce(self._flag)
        self.qv.multi_swap_peer(lposs, my_g, peers)
        dist.all_reduce(self._flag)
        self.bytes_exchanged += int((1 << self.nl) * (1 - 0.5 ** k)) * self.amps.element_size() * 2
        return 3

    def run_plan(self, plan, stats=None, queued=True):
        """Execute a plan; with `queued`, dense 1-/2-qubit gates between exchanges are flushed through the
        tile-blocked multi-gate passes (b200sv_apply_gate_sequence)."""
        from .executor import apply_ops_queued
        if queued:
            apply_ops_queued(self.qv, plan, stats, special={"swap": self.apply, "mswap": self.apply})
        else:
            for op in plan:
                n_launch = self.apply(op)
                if stats is not None:
                    stats["launches"] = stats.get("launches", 0) + n_launch

    def _staging(self, count):
        if self._tmp is None or self._tmp[0].shape[0] < count:
            self._tmp = [torch.empty((count, 2), dtype=self.amps.dtype, device=self.amps.device) for _ in range(2)]
        return self._tmp

    def swap_global(self, lpos, gbit):
        """Exchange local physical qubit `lpos` with global bit `gbit` (apply_chunk_swap,
        qubitvector.hpp:1753-1790): the rank whose global bit is 0 trades its lpos=1 half for the
        partner's lpos=0 half."""
        nl = self.nl
        peer = self.rank ^ (1 << gbit)
        upper = (self.rank >> gbit) & 1
        bit = 0 if upper else 1
        half = 1 << (nl - 1)
        launches = 0
        if self.exchange == "p2p":
            # In-place swap kernel over the peer mapping: each rank moves half of the pairs, so both NVLink
            # directions carry S*2^(nl-1) bytes and nothing is staged.  The tiny all-reduces are
            # stream-ordered rendezvous: the partner's earlier kernels are complete before its memory is
            # touched, and nobody runs ahead before the partner's half of the swap has landed.
            dist.all_reduce(self._flag)
            self.qv.chunk_swap_peer(lpos, self._peer_ptr[gbit], upper, upper)
            dist.all_reduce(self._flag)
            self.bytes_exchanged += half * self.amps.element_size() * 2
            return 3
        if lpos >= self.min_run_bits or (1 << lpos) >= self.slice_amps:
            c = min(self.slice_amps, 1 << lpos)
            tmp = self._staging(c)
            nslices = half // c
            pending = {}

            def region(i):
                start = _insert_zero(i * c, lpos) | (bit << lpos)
                return self.amps[start:start + c]

            def issue(i):
                ops = [dist.P2POp(dist.isend, region(i), peer), dist.P2POp(dist.irecv, tmp[i & 1][:c], peer)]
                if upper:
                    ops.reverse()
                pending[i] = dist.batch_isend_irecv(ops)

            for i in range(min(2, nslices)):
                issue(i)
            for i in range(nslices):
                for w in pending.pop(i):
                    w.wait()
                region(i).copy_(tmp[i & 1][:c])
                launches += 1
                if i + 2 < nslices:
                    issue(i + 2)
        else:  # low local qubit: gather the strided half through pack / unpack kernels
            c = min(self.slice_amps, half)
            stage = self._staging(2 * c)
            send = [stage[0][:c], stage[1][:c]]
            recv = [stage[0][c:2 * c], stage[1][c:2 * c]]
            nslices = half // c
            for i in range(nslices):
                s, r = send[i & 1], recv[i & 1]
                self.qv.pack_half(lpos, bit, i * c, c, s.data_ptr())
                ops = [dist.P2POp(dist.isend, s, peer), dist.P2POp(dist.irecv, r, peer)]
                if upper:
                    ops.reverse()
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                self.qv.unpack_half(lpos, bit, i * c, c, r.data_ptr())
                launches += 2
        self.bytes_exchanged += half * self.amps.element_size() * 2
        return launches

    # ---------------------------------------------------------------- reductions over ranks
    def _allreduce(self, arr):
        with self._ctx():
            t = torch.as_tensor(np.asarray(arr, dtype=np.float64)).to(self.amps.device)
            dist.all_reduce(t)
            return t.cpu().numpy()

    def norm(self):
        return float(self._allreduce([self.qv.norm()])[0])

    def expval_pauli(self, qubits, pauli):
        """Logical qubits; X/Y factors must be local (swap them in), Z factors on global qubits become a sign."""
        N = len(qubits)
        pq, pp = [], []
        sign = 1.0
        moved = False
        for i, q in enumerate(qubits):
            ch = pauli[N - 1 - i]
            if ch == "I":
                continue
            if self.phys[q] >= self.nl and ch in "XY":
                inv = {p: l for l, p in enumerate(self.phys)}
                busy = {qq for qq in qubits}
                victim = next(inv[p] for p in range(self.nl - 1, -1, -1) if inv[p] not in busy)
                lpos, gpos = self.phys[victim], self.phys[q]
                with self._ctx():
                    self.swap_global(lpos, gpos - self.nl)
                self.phys[victim], self.phys[q] = gpos, lpos
                moved = True
            p = self.phys[q]
            if p >= self.nl:  # Z on a global qubit: (-1)^bit of this rank
                if (self.rank >> (p - self.nl)) & 1:
                    sign = -sign
            else:
                pq.append(p)
                pp.append(ch)
        del moved
        if pq:
            local = self.qv.expval_pauli(pq, "".join(reversed(pp)))
        else:
            local = self.qv.norm()
        return float(self._allreduce([sign * local])[0])

    def restore_order(self):
        """Bring every logical qubit back to its own physical position (the reference's cache-blocking
        pass does the same before measure / save ops so that chunked and unchunked runs sample
        identically, cacheblocking.hpp `restore_qubit_map`; test/terra/backends/aer_simulator/
        test_chunk.py:31-168 asserts exact count equality).  Global positions first (exchanges),
        then the local permutation by transpositions (mcswap passes).  Returns launches issued."""
        nl, launches = self.nl, 0
        with self._ctx():
            for g in range(nl, self.n):
                inv = {p: l for l, p in enumerate(self.phys)}
                if inv[g] == g:
                    continue
                p = self.phys[g]            # where logical qubit g lives now
                if p >= nl:                 # on another global position: pull it to a local one first
                    lpos = nl - 1
                    victim = inv[lpos]
                    launches += self.swap_global(lpos, p - nl)
                    self.phys[victim], self.phys[g] = p, lpos
                    p = lpos
                    inv = {pp: l for l, pp in enumerate(self.phys)}
                occupant = inv[g]
                launches += self.swap_global(p, g - nl)
                self.phys[occupant], self.phys[g] = p, g
        for q in range(nl):
            p = self.phys[q]
            if p == q:
                continue
            inv = {pp: l for l, pp in enumerate(self.phys)}
            other = inv[q]
            self.qv.apply_mcswap([p, q])
            launches += 1
            self.phys[q], self.phys[other] = q, p
        assert self.phys == list(range(self.n))
        return launches

    def physical_to_logical(self, samples):
        """Map sampled physical basis-state indices to logical bit order."""
        s = np.asarray(samples, dtype=np.uint64)
        out = np.zeros_like(s)
        for q, p in enumerate(self.phys):
            out |= ((s >> np.uint64(p)) & np.uint64(1)) << np.uint64(q)
        return out

    def sample_measure(self, rnds, exact_order=True):
        """Executor::sample_measure (statevector_executor.hpp:1149-1227): route each draw to the rank
        whose cumulative-norm interval contains it, sample locally, combine.  With exact_order the
        qubit order is restored first so that the sampled indices equal the unsharded reference's
        for the same draws; otherwise sampling happens in physical order (same distribution, different
        draw -> outcome assignment) and the indices are mapped back bit by bit."""
        if exact_order:
            self.restore_order()
        rnds = np.asarray(rnds, dtype=np.float64)
        norms = np.zeros(self.world)
        norms[self.rank] = self.qv.norm()
        norms = self._allreduce(norms)
        cum = np.concatenate([[0.0], np.cumsum(norms)])
        lo, hi = cum[self.rank], cum[self.rank + 1]
        if self.rank == self.world - 1:
            mine = rnds >= lo
        else:
            mine = (rnds >= lo) & (rnds < hi)
        out = np.zeros(rnds.size, dtype=np.float64)  # exact for indices < 2^53
        if mine.any():
            local = self.qv.sample_measure(rnds[mine] - lo)
            out[mine] = (local + np.uint64(self.rank << self.nl)).astype(np.float64)
        out = self._allreduce(out)
        return self.physical_to_logical(out.astype(np.uint64))


# ======================================================================================================================
class ShardedState:
    """The C++ sharded executor (csrc/sharded.cu, b200sv_sharded_*): epoch planning, gate queues, tile passes, the
    staged / pipelined global-qubit exchange and the cross-shard reductions all run inside libb200sv.  This class only
    encodes op tuples and -- in the one-process-per-GPU layout -- moves 128-byte IPC blobs and per-rank scalars with
    torch.distributed (plumbing).

    Single process, several shards (Aer's Controller layout; device ids may repeat, which puts several shards on one
    GPU -- how the 1-GPU test tier covers the exchange paths):
        ShardedState(n, devices=[0, 1, 2, 3])
    One process per GPU under torchrun:
        ShardedState(n, world=W, rank=r, device=local_rank, dist=torch.distributed)
    """

    KIND = {"matrix": 0, "diagonal": 1, "mcx": 2, "mcy": 3, "mcphase": 4, "mcswap": 5, "mcu": 6}

    def __init__(self, n, devices=None, world=None, rank=None, device=0, dist=None, dtype=np.complex128,
                 staging_bytes=None):
        import ctypes as C
        from . import capi
        self._C, self._capi, self._lib = C, capi, capi.lib()
        self.n = int(n)
        self.dist = dist
        self.dtype = np.dtype(dtype)
        prec = 64 if self.dtype == np.complex128 else 32
        if devices is not None:
            self.world = len(devices)
            self.local_ranks = list(range(self.world))
            devs = [int(d) for d in devices]
        else:
            self.world, self.local_ranks, devs = int(world), [int(rank)], [int(device)]
        self.gbits = int(np.log2(self.world))
        self.nl = self.n - self.gbits
        self.h = C.c_void_p()
        lr = (C.c_int * len(devs))(*self.local_ranks)
        dv = (C.c_int * len(devs))(*devs)
        sb = (1 << 64) - 1 if staging_bytes is None else int(staging_bytes)
        capi.check(self._lib.b200sv_sharded_create(C.byref(self.h), self.n, prec, self.world, len(devs), lr, dv,
                                                   C.c_uint64(sb)))
        if devices is None and self.world > 1:
            blob = C.create_string_buffer(128)
            capi.check(self._lib.b200sv_sharded_ipc_export(self.h, self.local_ranks[0], blob))
            blobs = [None] * self.world
            dist.all_gather_object(blobs, bytes(blob.raw))
            for r in range(self.world):
                if r != self.local_ranks[0]:
                    capi.check(self._lib.b200sv_sharded_ipc_attach(self.h, r, C.create_string_buffer(blobs[r], 128)))
            dist.barrier()

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.b200sv_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- op encoding (the slice of Statevector::State::apply_op's gate table the workloads use,
    # statevector_state.hpp:314-384,757-960)
    @classmethod
    def encode(cls, ops):
        from .fusion import gate_matrix
        K = cls.KIND
        kinds, off, qs, doff, data = [], [0], [], [0], []

        def put(kind, qubits, payload=()):
            kinds.append(kind)
            qs.extend(int(q) for q in qubits)
            off.append(len(qs))
            data.extend(payload)
            doff.append(len(data))

        def flat(z):
            z = np.ascontiguousarray(np.asarray(z, dtype=np.complex128).reshape(-1))
            return z.view(np.float64).tolist()

        for op in ops:
            if op[0] == "unitary":
                put(K["matrix"], op[1], flat(np.asarray(op[2], dtype=np.complex128).reshape(-1, order="F")))
            elif op[0] == "diagonal":
                put(K["diagonal"], op[1], flat(op[2]))
            elif op[0] == "gate":
                name, qubits, params = op[1], op[2], op[3]
                if name in ("x", "cx", "ccx", "mcx"):
                    put(K["mcx"], qubits)
                elif name in ("y", "cy"):
                    put(K["mcy"], qubits)
                elif name in ("z", "cz", "ccz"):
                    put(K["mcphase"], qubits, [-1.0, 0.0])
                elif name in ("cp", "p", "mcphase", "cu1"):
                    ph = np.exp(1j * params[0])
                    put(K["mcphase"], qubits, [ph.real, ph.imag])
                elif name in ("s", "sdg", "t", "tdg"):
                    ph = {"s": 1j, "sdg": -1j, "t": np.exp(0.25j * np.pi), "tdg": np.exp(-0.25j * np.pi)}[name]
                    put(K["mcphase"], qubits, [ph.real, ph.imag])
                elif name in ("swap", "cswap"):
                    put(K["mcswap"], qubits)
                elif name == "rz":
                    put(K["diagonal"], qubits, flat(np.diag(gate_matrix("rz", params))))
                elif name in ("h", "sx"):
                    put(K["mcu"], qubits, flat(np.asarray(gate_matrix(name, params)).reshape(-1, order="F")))
                else:
                    raise ValueError("unsupported gate %s" % name)
            else:
                raise ValueError(op[0])
        return (np.asarray(kinds, dtype=np.int32), np.asarray(off, dtype=np.int32),
                np.asarray(qs if qs else [0], dtype=np.int32), np.asarray(doff, dtype=np.int64),
                np.asarray(data if data else [0.0], dtype=np.float64))

    @classmethod
    def plan_only(cls, n, world, ops, staging_bytes, dtype=np.complex128):
        """Host-only: passes / exchanges / overlap decisions apply_ops would take (no device)."""
        import ctypes as C
        from . import capi
        kinds, off, qs, doff, data = cls.encode(ops)
        out = np.zeros(8)
        capi.check(capi.lib().b200sv_sharded_plan_only(
            int(n), 64 if np.dtype(dtype) == np.complex128 else 32, int(world), C.c_uint64(int(staging_bytes)), int(kinds.size),
            kinds.ctypes.data_as(C.POINTER(C.c_int)), off.ctypes.data_as(C.POINTER(C.c_int)),
            qs.ctypes.data_as(C.POINTER(C.c_int)), doff.ctypes.data_as(C.POINTER(C.c_int64)),
            data.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double))))
        keys = ("passes", "exchanges", "staged", "inplace", "passes_overlapped", "slabs_max", "slabs_min", "qubit_swaps")
        return {k: int(v) for k, v in zip(keys, out)}

    @classmethod
    def selftest(cls, n, world, ops, staging_bytes, host_state):
        """Host-only interpretation of what apply_ops + restore_order would do (no device): `host_state` (2^n complex,
        logical order) is transformed in place; returns the run's statistics."""
        import ctypes as C
        from . import capi
        kinds, off, qs, doff, data = cls.encode(ops)
        if not host_state.flags["C_CONTIGUOUS"] or host_state.size != 1 << n:
            raise ValueError("host_state must be a contiguous array of 2^n amplitudes")
        prec = 64 if host_state.dtype == np.complex128 else 32
        out = np.zeros(8)
        capi.check(capi.lib().b200sv_sharded_selftest(
            int(n), prec, int(world), C.c_uint64(int(staging_bytes)), int(kinds.size),
            kinds.ctypes.data_as(C.POINTER(C.c_int)), off.ctypes.data_as(C.POINTER(C.c_int)),
            qs.ctypes.data_as(C.POINTER(C.c_int)), doff.ctypes.data_as(C.POINTER(C.c_int64)),
            data.ctypes.data_as(C.POINTER(C.c_double)), host_state.ctypes.data_as(C.c_void_p),
            out.ctypes.data_as(C.POINTER(C.c_double))))
        keys = ("passes", "exchanges", "staged", "inplace", "launches", "copies", "bytes_sent_per_shard", "overlapped_passes")
        return {k: (float(v) if k.startswith("bytes") else int(v)) for k, v in zip(keys, out)}

    # ---- execution
    def initialize(self):
        self._capi.check(self._lib.b200sv_sharded_initialize(self.h))

    def apply_ops(self, ops=None, encoded=None):
        C = self._C
        kinds, off, qs, doff, data = encoded if encoded is not None else self.encode(ops)
        self._capi.check(self._lib.b200sv_sharded_apply_ops(
            self.h, int(kinds.size), kinds.ctypes.data_as(C.POINTER(C.c_int)), off.ctypes.data_as(C.POINTER(C.c_int)),
            qs.ctypes.data_as(C.POINTER(C.c_int)), doff.ctypes.data_as(C.POINTER(C.c_int64)),
            data.ctypes.data_as(C.POINTER(C.c_double))))

    def synchronize(self):
        self._capi.check(self._lib.b200sv_sharded_synchronize(self.h))
        if self.dist is not None and self.world > 1 and len(self.local_ranks) < self.world:
            self.dist.barrier()

    def stats(self):
        out = np.zeros(8)
        self._capi.check(self._lib.b200sv_sharded_stats(self.h, out.ctypes.data_as(self._C.POINTER(self._C.c_double))))
        keys = ("passes", "exchanges", "staged", "inplace", "launches", "copies", "bytes_sent_per_shard", "overlapped_passes")
        return {k: (float(v) if k.startswith("bytes") else int(v)) for k, v in zip(keys, out)}

    def profile(self, on):
        self._capi.check(self._lib.b200sv_sharded_profile(self.h, 1 if on else 0))

    def profile_read(self):
        out = np.zeros(12)
        self._capi.check(self._lib.b200sv_sharded_profile_read(self.h, out.ctypes.data_as(self._C.POINTER(self._C.c_double))))
        keys = ("tile_pass", "exchange_region", "push", "unstage", "slab_pass", "exchange_inplace")
        return {k: {"count": int(out[2 * i]), "ms": float(out[2 * i + 1])} for i, k in enumerate(keys)}

    def compute_stream(self):
        """cudaStream_t of the first local shard's compute stream (for CUDA-event timing from the caller)."""
        C = self._C
        hh, p = C.c_void_p(), C.c_void_p()
        self._capi.check(self._lib.b200sv_sharded_shard_handle(self.h, self.local_ranks[0], C.byref(hh)))
        self._capi.check(self._lib.b200sv_stream(hh, C.byref(p)))
        return p.value

    def elapsed_ms(self):
        v = self._C.c_double(0)
        self._capi.check(self._lib.b200sv_sharded_elapsed_ms(self.h, self._C.byref(v)))
        return v.value

    @property
    def phys(self):
        a = np.zeros(self.n, dtype=np.int32)
        self._capi.check(self._lib.b200sv_sharded_qubit_map(self.h, a.ctypes.data_as(self._C.POINTER(self._C.c_int))))
        return [int(x) for x in a]

    def restore_order(self):
        self._capi.check(self._lib.b200sv_sharded_restore_order(self.h))

    # ---- reductions (per-process partials summed with torch.distributed when the shards live in several processes)
    def _sum(self, arr):
        arr = np.asarray(arr, dtype=np.float64)
        if self.dist is None or len(self.local_ranks) == self.world:
            return arr
        import torch
        t = torch.as_tensor(arr).cuda()
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def norms(self):
        self.synchronize()
        out = np.zeros(self.world)
        self._capi.check(self._lib.b200sv_sharded_norms(self.h, out.ctypes.data_as(self._C.POINTER(self._C.c_double))))
        return self._sum(out)

    def norm(self):
        return float(self.norms().sum())

    def expval_pauli(self, qubits, pauli):
        C = self._C
        self.synchronize()
        q = np.ascontiguousarray(list(qubits), dtype=np.uint64)
        v = C.c_double(0)
        self._capi.check(self._lib.b200sv_sharded_expval_pauli(self.h, q.ctypes.data_as(C.POINTER(C.c_uint64)), int(q.size),
                                                               pauli.encode(), C.byref(v)))
        return float(self._sum([v.value])[0])

    def sample_measure(self, rnds, exact_order=True):
        C = self._C
        if exact_order:
            self.restore_order()
        norms = self.norms()
        rnds = np.ascontiguousarray(rnds, dtype=np.float64)
        out = np.zeros(rnds.size, dtype=np.uint64)
        self._capi.check(self._lib.b200sv_sharded_sample_measure(
            self.h, rnds.ctypes.data_as(C.POINTER(C.c_double)), int(rnds.size), norms.ctypes.data_as(C.POINTER(C.c_double)),
            out.ctypes.data_as(C.POINTER(C.c_uint64))))
        if self.dist is not None and len(self.local_ranks) < self.world:
            return self._sum(out.astype(np.float64)).astype(np.uint64)  # exact below 2^53
        return out

    def shard_vector(self, rank):
        """Amplitudes of a local shard in PHYSICAL order (call restore_order() first for the logical slice)."""
        C = self._C
        self.synchronize()
        hh = C.c_void_p()
        self._capi.check(self._lib.b200sv_sharded_shard_handle(self.h, int(rank), C.byref(hh)))
        out = np.empty(1 << self.nl, dtype=self.dtype)
        self._capi.check(self._lib.b200sv_download(hh, out.ctypes.data_as(C.c_void_p), 0, 1 << self.nl))
        return out

    def vector(self):
        """The whole register in logical order (single-process layout, small n: tests)."""
        self.restore_order()
        return np.concatenate([self.shard_vector(r) for r in self.local_ranks])

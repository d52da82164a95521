"""TEST-SIDE host model of a state sharded over 2^g ranks by *global qubits* (torch.distributed over gloo,
shards driven by the CPU checker).  The product is csrc/sharded.cu behind the C ABI (qb200_sv_*); this file
restates its host logic -- qubit map, matrix re-indexing, global controls, the index arithmetic of both
exchange kernels -- so that world_size 2/4 CPU tests can replay the schedules of the library's planner
(qb200_sv_plan) and compare with an unsharded run.  Round 1 shipped this as qsim_b200/sharded.py.

Reference precedent: the cuStateVecEx backend's wire ordering + index-bit swaps
(lib/vectorspace_custatevecex.h:163-177, lib/simulator_custatevecex.h:147-196,
lib/run_custatevecex.h:243-305).  The policy below is ours (the swaps of the
reference happen inside the closed cuStateVecEx library).

Layout.  Physical index bit p of an amplitude: p < n_local addresses the
amplitude inside the rank's shard, p >= n_local is bit (p - n_local) of the rank.
`pos[q]` is the physical bit that currently holds logical qubit q.  A gate whose
qubits all sit on local bits is the single-GPU kernel on every shard (matrix
re-indexed when the physical order differs from the logical one).  A gate that
touches a global qubit first triggers a local<->global swap:

  1. (local)    victims -- chosen by furthest next use over the remaining gate
                list -- are moved to the TOP local bits with 2-qubit SWAP passes
                (HBM-speed, 16*2^n_local bytes each);
  2. (exchange) the top k local bits are exchanged with k rank bits: the shard
                splits into 2^k contiguous slices, slice b goes to the peer
                whose k rank bits equal b and is replaced by that peer's slice
                (grouped send/recv inside each 2^k-rank group, staged through a
                bounded transfer buffer).  Bytes sent = bytes received =
                shard * (1 - 2^-k) per rank.

When a swap is forced, ALL global qubits that are not among the g furthest-used
qubits are exchanged in the same step (cost grows only as 1 - 2^-k), which is
what keeps the swap count low on RQCs (see plan_swaps()).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------------------
# planning (pure host logic, no device, no torch)
# ---------------------------------------------------------------------------
@dataclass
class SwapStep:
    """Before op `before_op`: bring `victims` (logical qubits, currently local) to the
    top local bits, then exchange them with `incoming` (logical qubits, currently global)."""
    before_op: int
    victims: List[int]
    incoming: List[int]


def next_use_table(op_qubits: Sequence[Sequence[int]], num_qubits: int) -> List[Dict[int, int]]:
    """next_use[i][q] = index of the first op >= i that touches q (len(ops) if none)."""
    nxt = {q: len(op_qubits) for q in range(num_qubits)}
    table = [None] * (len(op_qubits) + 1)
    table[len(op_qubits)] = dict(nxt)
    for i in range(len(op_qubits) - 1, -1, -1):
        for q in op_qubits[i]:
            nxt[q] = i
        table[i] = dict(nxt)
    return table


def plan_swaps(op_qubits: Sequence[Sequence[int]], num_qubits: int, num_global: int,
               initial_global: Optional[Sequence[int]] = None) -> List[SwapStep]:
    """Look-ahead remap pass: Belady-style choice of the global set.

    op_qubits[i] = all qubits (targets and controls) op i touches.  Returns the swap
    steps; between them every op touches local qubits only."""
    g = num_global
    if g == 0:
        return []
    glob = list(initial_global) if initial_global is not None else list(range(num_qubits - g, num_qubits))
    assert len(glob) == g
    nxt = next_use_table(op_qubits, num_qubits)
    steps = []
    for i, qs in enumerate(op_qubits):
        if len(qs) > num_qubits - g:
            raise ValueError("gate touches more qubits than a shard holds")
        if not any(q in glob for q in qs):
            continue
        # desired global set: the g qubits whose next use (from op i on) is furthest,
        # never a qubit of this op; ties broken towards keeping current globals global
        cand = [q for q in range(num_qubits) if q not in qs]
        cand.sort(key=lambda q: (-nxt[i][q], 0 if q in glob else 1, q))
        want = set(cand[:g])
        incoming = [q for q in glob if q not in want]           # leave the global set
        victims = [q for q in sorted(want) if q not in glob]    # enter the global set
        assert len(incoming) == len(victims) and incoming
        steps.append(SwapStep(i, victims, incoming))
        glob = [q for q in glob if q in want] + victims
    return steps


def reindex_matrix(matrix: np.ndarray, phys: Sequence[int]) -> Tuple[List[int], np.ndarray]:
    """Gate matrix given for qubits at physical bits `phys` (bit k of the matrix index
    <-> phys[k], not necessarily ascending) -> (sorted bits, matrix for sorted order)."""
    g = len(phys)
    order = sorted(range(g), key=lambda k: phys[k])
    if order == list(range(g)):
        return list(phys), matrix
    dim = 1 << g
    m = np.asarray(matrix).reshape(dim, dim, 2) if not np.iscomplexobj(matrix) else np.asarray(matrix).reshape(dim, dim)
    idx = np.zeros(dim, dtype=np.int64)  # idx[new] = old
    for new in range(dim):
        old = 0
        for j, k in enumerate(order):
            old |= ((new >> j) & 1) << k
        idx[new] = old
    out = m[np.ix_(idx, idx)]
    return [phys[k] for k in order], np.ascontiguousarray(out).reshape(-1) if not np.iscomplexobj(matrix) else np.ascontiguousarray(out)


SWAP_MATRIX = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex64)


# ---------------------------------------------------------------------------
# the sharded simulator (engine-agnostic; the product engine is B200Engine)
# ---------------------------------------------------------------------------
class B200Engine:
    """Local shard on one B200: torch CUDA tensor for the bytes (so torch.distributed
    can move slices of it), libqsim_b200 kernels for everything else."""

    def __init__(self, n_local: int, device_index: int, dtype=np.float32, p2p: bool = False):
        """p2p=False: the shard is a torch tensor and swaps go through torch.distributed
        (grouped NCCL send/recv, staged).  p2p=True: the shard is a cudaMalloc'ed buffer
        exported with CUDA IPC and swaps are ONE kernel per GPU over NVLink peer memory
        (qb200_swap_global_local), in place: no staging buffer, no local SWAP passes."""
        import torch
        from . import backend
        self.torch = torch
        self.n_local = n_local
        self.device = torch.device("cuda", device_index)
        self.p2p = p2p
        self.element_size = np.dtype(dtype).itemsize
        self.ss = backend.StateSpaceB200(dtype, device=device_index)
        self.sim = backend.SimulatorB200(dtype, device=device_index)
        if p2p:
            self.shard = None
            self.state = self.ss.Create(n_local)
            if self.ss.IsNull(self.state):
                raise MemoryError("not enough device memory for the shard")
        else:
            tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
            self.shard = torch.empty(2 << n_local, dtype=tdt, device=self.device)
            self.state = self.ss.CreateFromPointer(self.shard.data_ptr(), n_local)
        self._stage = None
        self._peers = None
        self._flag = None

    # ---- NVLink peer-memory path ------------------------------------------------
    def connect_peers(self, dist, rank: int, world: int):
        """exchange CUDA IPC handles of all shards (once)."""
        import ctypes as C
        lib = self.ss._lib
        h = (C.c_ubyte * 64)()
        self.ss._check(lib.qb200_ipc_export(self.state.get(), h), "ipc_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h))
        self._peers = [None] * world
        for r in range(world):
            if r == rank:
                continue
            hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
            ptr = C.c_void_p()
            self.ss._check(lib.qb200_ipc_import(hb, C.byref(ptr)), "ipc_import")
            self._peers[r] = ptr.value
        self._flag = self.torch.zeros(1, device=self.device)

    def stream_barrier(self, dist):
        """stream-ordered barrier: no rank's later kernels start before every rank's earlier
        kernels are done (tiny all-reduce on the kernels' stream; the host does not block)."""
        dist.all_reduce(self._flag)

    def swap_global_local(self, peer_ranks, k, local_bits, my_value):
        import ctypes as C
        arr = (C.c_void_p * (1 << k))(*[self._peers[r] if r is not None else None for r in peer_ranks])
        lb = (C.c_uint * k)(*local_bits)
        self.sim._check(self.sim._lib.qb200_swap_global_local(self.sim._ctx, self.sim._dt, self.state.get(),
                                                              self.n_local, arr, k, lb, my_value), "swap_global_local")

    def zero(self):
        self.ss.SetAllZeros(self.state)

    def set_ampl(self, i, val):
        self.ss.SetAmpl(self.state, i, val)

    def get_ampl(self, i):
        return self.ss.GetAmpl(self.state, i)

    def apply_gate(self, qs, matrix):
        self.sim.ApplyGate(qs, matrix, self.state)

    def apply_controlled_gate(self, qs, cqs, cvals, matrix):
        self.sim.ApplyControlledGate(qs, cqs, cvals, matrix, self.state)

    def norm(self) -> float:
        return self.ss.Norm(self.state)

    def expectation_value(self, qs, matrix) -> complex:
        return self.sim.ExpectationValue(qs, matrix, self.state)

    def slice(self, start_scalar: int, count_scalar: int):
        return self.shard[start_scalar:start_scalar + count_scalar]

    def staging(self, count_scalar: int):
        if self._stage is None or self._stage.numel() < count_scalar:
            self._stage = self.torch.empty(count_scalar, dtype=self.shard.dtype, device=self.device)
        return self._stage[:count_scalar]

    def sync(self):
        self.torch.cuda.synchronize(self.device)

    def event(self):
        """CUDA event recorded on the stream the kernels and the collectives are ordered on."""
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def to_numpy(self):
        return self.ss.to_numpy(self.state)


@dataclass
class SwapStats:
    swaps: int = 0
    local_swap_passes: int = 0
    bytes_sent: int = 0          # per rank
    exchange_seconds: float = 0.0
    detail: List[Tuple[int, int, int]] = field(default_factory=list)  # (before_op, k, bytes)


class ShardedSimulator:
    """n-qubit state over world_size = 2^g ranks.  `engine` holds the local shard,
    `dist` is torch.distributed (already initialised) or None for a single rank."""

    MIN_SWAP_BIT = 4  # victims below this local bit are moved up before the peer-memory exchange

    def __init__(self, num_qubits: int, engine, dist=None, rank: int = 0, world_size: int = 1,
                 transfer_scalars: int = 1 << 28):
        g = world_size.bit_length() - 1
        if 1 << g != world_size:
            raise ValueError("world_size must be a power of two")
        if g + 2 > num_qubits:
            # same guard as lib/multiprocess_custatevecex.h:160-163
            raise ValueError("too few qubits to shard over this many ranks")
        self.n, self.g, self.n_local = num_qubits, g, num_qubits - g
        self.engine, self.dist, self.rank, self.world = engine, dist, rank, world_size
        self.pos = list(range(num_qubits))       # logical qubit -> physical bit
        self.transfer_scalars = transfer_scalars  # staging buffer bound per peer piece
        self.stats = SwapStats()
        self._exchange_events = []

    # ---- bookkeeping -------------------------------------------------------
    def global_qubits(self) -> List[int]:
        return [q for q in range(self.n) if self.pos[q] >= self.n_local]

    def qubit_at(self, phys: int) -> int:
        return self.pos.index(phys)

    def set_state_zero(self):
        self.engine.zero()
        if self.rank == 0:
            self.engine.set_ampl(0, 1.0)

    # ---- gates ---------------------------------------------------------------
    def _local_gate(self, qs, cqs, cvals, matrix):
        phys = [self.pos[q] for q in qs]
        sorted_bits, m = reindex_matrix(matrix, phys)
        if cqs:
            # controls: split into local controls (passed on) and global controls
            # (decided per rank: the whole shard either qualifies or is skipped)
            order = sorted(range(len(cqs)), key=lambda k: cqs[k])  # cvals bit i <-> i-th lowest control
            loc, loc_vals, ok = [], [], True
            for i, k in enumerate(order):
                p, v = self.pos[cqs[k]], (cvals >> i) & 1
                if p >= self.n_local:
                    ok &= ((self.rank >> (p - self.n_local)) & 1) == v
                else:
                    loc.append(p); loc_vals.append(v)
            if not ok:
                return
            if loc:
                o2 = sorted(range(len(loc)), key=lambda k: loc[k])
                cv = sum(loc_vals[k] << i for i, k in enumerate(o2))
                self.engine.apply_controlled_gate(sorted_bits, [loc[k] for k in o2], cv, m)
                return
        self.engine.apply_gate(sorted_bits, m)

    def apply_gate(self, qs, matrix, cqs=(), cvals=0):
        """Applies one gate; every target must currently be local (run() guarantees it)."""
        if any(self.pos[q] >= self.n_local for q in qs):
            raise RuntimeError("gate target on a global qubit: call run() or swap first")
        self._local_gate(list(qs), list(cqs), cvals, matrix)

    # ---- swaps -----------------------------------------------------------------
    def _local_swap(self, pa: int, pb: int):
        """exchange two local physical bits with one 2-qubit SWAP pass."""
        if pa == pb:
            return
        lo, hi = min(pa, pb), max(pa, pb)
        self.engine.apply_gate([lo, hi], SWAP_MATRIX)
        qa, qb = self.qubit_at(pa), self.qubit_at(pb)
        self.pos[qa], self.pos[qb] = pb, pa
        self.stats.local_swap_passes += 1

    def swap(self, victims: Sequence[int], incoming: Sequence[int], before_op: int = -1):
        """victims: local logical qubits that become global; incoming: global ones that
        become local.  len(victims) == len(incoming) == k."""
        import time
        k = len(victims)
        assert k == len(incoming) and k >= 1
        if getattr(self.engine, "remap", False) and self.dist is not None and self.world > 1:
            return self._swap_remap(list(victims), list(incoming), before_op)
        if getattr(self.engine, "p2p", False) and self.dist is not None and self.world > 1:
            return self._swap_p2p(victims, incoming, before_op)
        top = list(range(self.n_local - k, self.n_local))
        # 1. bring the victims to the top k local bits (skip those already there)
        need = [self.pos[v] for v in victims if self.pos[v] not in top]
        free_top = [p for p in top if self.qubit_at(p) not in victims]
        for src, dst in zip(sorted(need), free_top):
            self._local_swap(src, dst)
        # 2. exchange top local bit (n_local-k+j) with the rank bit that holds incoming[j]
        top_q = [self.qubit_at(p) for p in top]                # victim at each top bit
        gbits = [self.pos[q] - self.n_local for q in incoming]  # rank bit of each incoming qubit
        slice_scalars = (2 << self.n_local) >> k
        my = sum(((self.rank >> gb) & 1) << j for j, gb in enumerate(gbits))
        t0 = time.perf_counter()
        ev0 = self.engine.event() if hasattr(self.engine, "event") else None
        if self.dist is not None and self.world > 1:
            self._exchange(k, gbits, my, slice_scalars)
        if ev0 is not None:
            self._exchange_events.append((ev0, self.engine.event()))
        self.stats.exchange_seconds += time.perf_counter() - t0
        for j in range(k):
            self.pos[top_q[j]], self.pos[incoming[j]] = self.n_local + gbits[j], top[j]
        sent = (slice_scalars * ((1 << k) - 1)) * self.engine.shard.element_size() if getattr(self.engine, "shard", None) is not None else 0
        self.stats.swaps += 1
        self.stats.bytes_sent += sent
        self.stats.detail.append((before_op, k, sent))

    def _swap_p2p(self, victims, incoming, before_op):
        """in-place exchange over NVLink peer memory: the victims' local bits (wherever they
        are) are swapped with the incoming qubits' rank bits by one kernel per GPU."""
        k = len(victims)
        # The kernel moves 16-byte items; a swapped local bit b makes contiguous runs of 2^b
        # amplitudes, and NVLink wants >= 256-byte runs: 695 GB/s per direction for b >= 5,
        # 503 at b = 3, 392 at b = 2 (tools/swap_bench.py).  A victim sitting on one of the
        # lowest bits is first moved up by one local SWAP pass (2.5 ms per 8 GiB shard).
        taken = {self.pos[v] for v in victims}
        for v in victims:
            if self.pos[v] < self.MIN_SWAP_BIT:
                dst = next(p for p in range(self.n_local - 1, self.MIN_SWAP_BIT - 1, -1) if p not in taken)
                taken.discard(self.pos[v])
                self._local_swap(self.pos[v], dst)
                taken.add(dst)
        order = sorted(range(k), key=lambda j: self.pos[victims[j]])
        victims = [victims[j] for j in order]
        lbits = [self.pos[v] for v in victims]
        gbits = [self.pos[q] - self.n_local for q in incoming]
        my = sum(((self.rank >> gb) & 1) << j for j, gb in enumerate(gbits))
        peers = [None if b == my else self._peer(gbits, b) for b in range(1 << k)]
        eng = self.engine
        ev0 = eng.event()
        eng.stream_barrier(self.dist)
        eng.swap_global_local(peers, k, lbits, my)
        eng.stream_barrier(self.dist)
        self._exchange_events.append((ev0, eng.event()))
        for j in range(k):
            self.pos[victims[j]], self.pos[incoming[j]] = self.n_local + gbits[j], lbits[j]
        sent = ((2 << self.n_local) >> k) * ((1 << k) - 1) * eng.element_size
        self.stats.swaps += 1
        self.stats.bytes_sent += sent
        self.stats.detail.append((before_op, k, sent))

    def _swap_remap(self, victims, incoming, before_op):
        """out-of-place exchange, the index arithmetic of k_remap_push (csrc/sharded.cu): the victims' bits are
        squeezed out of the local index (order of the rest kept), every shard lands in slice `my` of each
        destination, incoming qubit j ends on local bit n_local - k + j."""
        k = len(victims)
        victims = sorted(victims, key=lambda q: self.pos[q])
        lbits = [self.pos[v] for v in victims]
        gbits = [self.pos[q] - self.n_local for q in incoming]
        my = sum(((self.rank >> gb) & 1) << j for j, gb in enumerate(gbits))
        dst_ranks = [self._peer(gbits, b) for b in range(1 << k)]
        self.engine.remap_push(dst_ranks, k, lbits, my)
        for q in range(self.n):
            p = self.pos[q]
            if p >= self.n_local or q in victims:
                continue
            self.pos[q] = p - sum(1 for b in lbits if b < p)
        for j in range(k):
            self.pos[victims[j]] = self.n_local + gbits[j]
            self.pos[incoming[j]] = self.n_local - k + j
        sent = ((2 << self.n_local) >> k) * ((1 << k) - 1) * self.engine.element_size
        self.stats.swaps += 1
        self.stats.bytes_sent += sent
        self.stats.detail.append((before_op, k, sent))

    def run_schedule(self, ops, schedule):
        """schedule: qsim_b200.sv.plan() output -- ("gate", i) / ("swap", victims, incoming)."""
        for step in schedule:
            if step[0] == "gate":
                op = ops[step[1]]
                self._local_gate(list(op.qubits), list(op.controls), op.cvals, op.matrix)
            else:
                self.swap(step[1], step[2])

    def _peer(self, gbits, b):
        r = self.rank
        for j, gb in enumerate(gbits):
            r = (r & ~(1 << gb)) | (((b >> j) & 1) << gb)
        return r

    def _exchange(self, k, gbits, my, slice_scalars):
        """slice b of the shard <-> slice `my` of the peer whose selected rank bits equal b.
        Staged through a bounded buffer: pieces of at most transfer_scalars per peer."""
        dist = self.dist
        peers = [(b, self._peer(gbits, b)) for b in range(1 << k) if b != my]
        piece = min(slice_scalars, self.transfer_scalars)
        for off in range(0, slice_scalars, piece):
            cnt = min(piece, slice_scalars - off)
            stage = self.engine.staging(cnt * len(peers))
            ops = []
            for i, (b, peer) in enumerate(peers):
                send = self.engine.slice(b * slice_scalars + off, cnt)
                recv = stage[i * cnt:(i + 1) * cnt]
                ops.append(dist.P2POp(dist.isend, send, peer))
                ops.append(dist.P2POp(dist.irecv, recv, peer))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for i, (b, peer) in enumerate(peers):
                self.engine.slice(b * slice_scalars + off, cnt).copy_(stage[i * cnt:(i + 1) * cnt])

    def exchange_device_ms(self) -> float:
        """device time of all exchanges so far (CUDA events; synchronises)."""
        if not self._exchange_events:
            return 0.0
        self.engine.sync()
        return float(sum(a.elapsed_time(b) for a, b in self._exchange_events))

    def reset_stats(self):
        self.stats = SwapStats()
        self._exchange_events = []

    # ---- whole circuits --------------------------------------------------------
    def run(self, ops, plan: Optional[List[SwapStep]] = None):
        """ops: objects with .qubits, .controls, .cvals, .matrix (qsim_b200.trace.TraceOp)."""
        if plan is None:
            plan = plan_swaps([list(o.qubits) + list(o.controls) for o in ops], self.n, self.g,
                              self.global_qubits())
        steps = {s.before_op: s for s in plan}
        for i, op in enumerate(ops):
            if i in steps:
                self.swap(steps[i].victims, steps[i].incoming, before_op=i)
            self._local_gate(list(op.qubits), list(op.controls), op.cvals, op.matrix)
        return plan

    # ---- reductions / access ------------------------------------------------------
    def norm(self) -> float:
        v = self.engine.norm()
        if self.dist is not None and self.world > 1:
            import torch
            t = torch.tensor([v], dtype=torch.float64, device=getattr(self.engine, "device", "cpu"))
            self.dist.all_reduce(t)
            v = float(t.item())
        return v

    def make_local(self, qubits: Sequence[int]):
        """Collective: one swap that brings every global qubit of `qubits` into the local part; the victims are
        the local qubits outside `qubits` on the highest physical bits."""
        qubits = list(qubits)
        incoming = [q for q in qubits if self.pos[q] >= self.n_local]
        if not incoming:
            return
        spare = sorted((q for q in range(self.n) if self.pos[q] < self.n_local and q not in qubits),
                       key=lambda q: -self.pos[q])
        if len(spare) < len(incoming):
            raise ValueError("not enough local qubits outside the operator to swap with")
        self.swap(spare[:len(incoming)], incoming)

    def expectation_value(self, qs: Sequence[int], matrix) -> complex:
        """<psi|M|psi> on a sharded state (SimulatorCUDA::ExpectationValue, lib/simulator_cuda.h:216-260): the
        operator's qubits are made local if they are not (one exchange), every rank runs the
        read-only pass on its shard and the P partial values are added.  Collective: every rank gets the value."""
        qs = list(qs)
        self.make_local(qs)
        sorted_bits, m = reindex_matrix(matrix, [self.pos[q] for q in qs])
        v = complex(self.engine.expectation_value(sorted_bits, m))
        if self.dist is not None and self.world > 1:
            import torch
            t = torch.tensor([v.real, v.imag], dtype=torch.float64, device=getattr(self.engine, "device", "cpu"))
            self.dist.all_reduce(t)
            v = complex(t[0].item(), t[1].item())
        return v

    def locate(self, logical_index: int) -> Tuple[int, int]:
        """logical amplitude index -> (rank, local index) under the current qubit map."""
        phys = 0
        for q in range(self.n):
            phys |= ((logical_index >> q) & 1) << self.pos[q]
        return phys >> self.n_local, phys & ((1 << self.n_local) - 1)

    def get_ampl(self, logical_index: int) -> complex:
        """collective: every rank gets the amplitude."""
        r, li = self.locate(logical_index)
        val = self.engine.get_ampl(li) if r == self.rank else 0j
        if self.dist is not None and self.world > 1:
            import torch
            t = torch.tensor([val.real, val.imag], dtype=torch.float64, device=getattr(self.engine, "device", "cpu"))
            self.dist.all_reduce(t)
            val = complex(t[0].item(), t[1].item())
        return val

"""Sharded state (global-qubit swaps as grouped send/recv) on CPU: world_size 2 and 4
over the gloo backend, local shards driven by the CPU oracle (tests only).  Checks the
host logic of the sharded state -- the library's swap planner (qb200_sv_plan, host-only C++), matrix
re-indexing, qubit map, the index arithmetic of both exchange kernels, global controls -- replayed by
tests/sharded_host_model.py against an unsharded oracle run of the same circuit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sharded_host_model import ShardedSimulator, SwapStep, plan_swaps, reindex_matrix  # noqa: E402
from qsim_b200.trace import TraceOp, read_trace  # noqa: E402


class OracleEngine:
    """Test-only local engine: numpy shard + CPU oracle kernels."""

    def __init__(self, n_local):
        import torch
        from oracle.oracle import Oracle
        self.orc = Oracle()
        self.n_local = n_local
        self.shard = torch.zeros(2 << n_local, dtype=torch.float32)
        self.np = self.shard.numpy().view(np.complex64)
        self.device = "cpu"
        self._stage = None

    def zero(self):
        self.np[:] = 0

    def set_ampl(self, i, val):
        self.np[i] = val

    def get_ampl(self, i):
        return complex(self.np[i])

    def apply_gate(self, qs, matrix):
        self.orc.apply_gate(self.np, qs, matrix)

    def apply_controlled_gate(self, qs, cqs, cvals, matrix):
        self.orc.apply_controlled_gate(self.np, qs, cqs, cvals, matrix)

    def norm(self):
        return self.orc.norm(self.np)

    def expectation_value(self, qs, matrix):
        return self.orc.expectation_value(self.np, qs, matrix)

    def slice(self, start, count):
        return self.shard[start:start + count]

    def staging(self, count):
        import torch
        if self._stage is None or self._stage.numel() < count:
            self._stage = torch.empty(count, dtype=torch.float32)
        return self._stage[:count]


class PeerOracleEngine(OracleEngine):
    """Emulates the NVLink peer-memory engine (B200Engine(p2p=True)) over gloo: the same in-place
    semantics as csrc/p2p_swap.cu -- the k local bits `local_bits` are swapped with k rank bits, for
    every peer value b != my the amplitudes whose local bits equal b trade places with the peer's
    amplitudes whose local bits equal my.  Exercises ShardedSimulator._swap_p2p on CPU."""
    p2p = True
    element_size = 4

    class _Event:
        def elapsed_time(self, other):
            return 0.0

    def __init__(self, n_local, dist):
        super().__init__(n_local)
        self.dist = dist

    def event(self):
        return self._Event()

    def stream_barrier(self, dist):
        dist.barrier()

    def swap_global_local(self, peer_ranks, k, local_bits, my_value):
        import torch
        assert list(local_bits) == sorted(local_bits) and len(local_bits) == k
        idx = np.arange(1 << self.n_local, dtype=np.int64)
        val = np.zeros_like(idx)
        for j, lb in enumerate(local_bits):
            val |= ((idx >> lb) & 1) << j
        reqs, incoming = [], {}
        for b in range(1 << k):
            if b == my_value:
                continue
            sel = np.nonzero(val == b)[0]
            out = torch.from_numpy(np.ascontiguousarray(self.np[sel]).view(np.float32).copy())
            inc = torch.empty_like(out)
            incoming[b] = (sel, inc)
            reqs.append(self.dist.isend(out, dst=peer_ranks[b]))
            reqs.append(self.dist.irecv(inc, src=peer_ranks[b]))
        for r in reqs:
            r.wait()
        for b, (sel, inc) in incoming.items():
            self.np[sel] = inc.numpy().view(np.complex64)


class RemapOracleEngine(PeerOracleEngine):
    """Emulates the out-of-place push exchange (k_remap_push, csrc/sharded.cu) over gloo."""
    remap = True

    def remap_push(self, dst_ranks, k, lbits, my):
        import torch
        nl = self.n_local
        idx = np.arange(1 << nl, dtype=np.int64)
        v = np.zeros_like(idx)
        rest = idx.copy()
        for j in range(k - 1, -1, -1):
            b = lbits[j]
            v |= ((rest >> b) & 1) << j
            rest = ((rest >> (b + 1)) << b) | (rest & ((1 << b) - 1))
        dest_idx = rest | (my << (nl - k))
        new = np.zeros_like(self.np)
        reqs, incoming = [], []
        for val in range(1 << k):
            sel = np.nonzero(v == val)[0]
            order = np.argsort(dest_idx[sel])
            payload = np.ascontiguousarray(self.np[sel][order])
            where = dest_idx[sel][order]            # the slice `my` of the destination, ascending
            if dst_ranks[val] == self.dist.get_rank():
                new[where] = payload
                continue
            out = torch.from_numpy(payload.view(np.float32).copy())
            inc = torch.empty_like(out)
            reqs.append(self.dist.isend(out, dst=dst_ranks[val]))
            reqs.append(self.dist.irecv(inc, src=dst_ranks[val]))
            incoming.append((val, inc))
        for r in reqs:
            r.wait()
        for val, inc in incoming:
            # the sender whose exchanged rank bits equal `val` fills slice `val` of my new buffer
            lo = val << (nl - k)
            new[lo:lo + (1 << (nl - k))] = inc.numpy().view(np.complex64)
        self.np[:] = new


def random_ops(n, count, seed, max_local):
    rs = np.random.RandomState(seed)
    ops = []
    for i in range(count):
        g = int(rs.randint(1, min(4, max_local) + 1))
        qs = sorted(rs.choice(n, g, replace=False).tolist())
        m = (rs.standard_normal((1 << g, 1 << g)) + 1j * rs.standard_normal((1 << g, 1 << g))).astype(np.complex64)
        u, _ = np.linalg.qr(m)
        cs, cv = [], 0
        if i % 4 == 3 and g <= 2:
            free = [q for q in range(n) if q not in qs]
            cs = sorted(rs.choice(free, int(rs.randint(1, 3)), replace=False).tolist())
            cv = int(rs.randint(0, 1 << len(cs)))
        ops.append(TraceOp(qs, cs, cv, u.astype(np.complex64).reshape(-1).view(np.float32).copy()))
    return ops


def expectation_cases(n):
    """(qubits, matrix) pairs: low / high (global after most plans) / mixed targets, 1 to 3 qubits."""
    rng = np.random.RandomState(77)
    out = []
    for qs in ([0], [n - 1], [1, n - 2], [n - 2, n - 1], [0, 3, n - 1], [2, 4, 5]):
        d = 1 << len(qs)
        out.append((qs, (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))).astype(np.complex64)))
    return out


def oracle_full(n, ops):
    from oracle.oracle import Oracle
    orc = Oracle()
    st = np.zeros(1 << n, np.complex64)
    st[0] = 1
    for op in ops:
        if op.controls:
            orc.apply_controlled_gate(st, op.qubits, op.controls, op.cvals, op.matrix)
        else:
            orc.apply_gate(st, op.qubits, op.matrix)
    return st


def _worker(rank, world, port, n, ops, transfer_scalars, out_dir, p2p=False, schedule=None):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = world.bit_length() - 1
        eng = RemapOracleEngine(n - g, dist) if p2p == "remap" else PeerOracleEngine(n - g, dist) if p2p else OracleEngine(n - g)
        sim = ShardedSimulator(n, eng, dist=dist, rank=rank, world_size=world, transfer_scalars=transfer_scalars)
        sim.set_state_zero()
        if schedule is not None:
            sim.run_schedule(ops, schedule)
            plan = [s for s in schedule if s[0] == "swap"]
        else:
            plan = sim.run(ops)
        norm = sim.norm()
        amp5 = sim.get_ampl(5)
        swaps, sent, lsp = sim.stats.swaps, sim.stats.bytes_sent, sim.stats.local_swap_passes
        # expectation values on the sharded state (operators of expectation_cases(n)): targets that sit on rank
        # bits are swapped in first, so the qubit map -- saved below -- may change, the logical state must not
        evs = [sim.expectation_value(qs, m) for qs, m in expectation_cases(n)]
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), shard=eng.np, pos=np.array(sim.pos), norm=norm,
                 amp5=np.array([amp5.real, amp5.imag]), swaps=swaps, nplan=len(plan),
                 bytes_sent=sent, local_swap_passes=lsp, evs=np.array(evs), swaps_after=sim.stats.swaps)
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_sharded(world, n, ops, tmp_path, transfer_scalars=1 << 28, p2p=False, schedule=None):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, free_port(), n, ops, transfer_scalars, str(tmp_path), p2p, schedule), nprocs=world, join=True)
    g = world.bit_length() - 1
    n_local = n - g
    full = np.zeros(1 << n, np.complex64)
    res = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    pos = res[0]["pos"]
    idx = np.arange(1 << n, dtype=np.int64)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> q) & 1) << int(pos[q])
    for r in range(world):
        assert np.array_equal(res[r]["pos"], pos)
        sel = (phys >> n_local) == r
        full[sel] = res[r]["shard"][phys[sel] & ((1 << n_local) - 1)]
    return full, res


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_random_circuit_matches_unsharded_oracle(world, tmp_path):
    n = 9
    g = world.bit_length() - 1
    ops = random_ops(n, 40, seed=world, max_local=n - g)
    want = oracle_full(n, ops)
    got, res = run_sharded(world, n, ops, tmp_path, transfer_scalars=64)  # tiny staging: forces chunking
    assert np.abs(got - want).max() < 2e-6
    assert abs(float(res[0]["norm"]) - 1.0) < 1e-5
    a5 = res[0]["amp5"]
    assert abs(complex(a5[0], a5[1]) - want[5]) < 2e-6
    assert int(res[0]["swaps"]) >= 1 and int(res[0]["swaps"]) == int(res[0]["nplan"])
    check_expectations(n, want, res)


def check_expectations(n, want, res):
    """every rank holds the same values, equal to the oracle's on the unsharded state; at least one operator
    needed its qubits swapped in."""
    from oracle.oracle import Oracle
    orc = Oracle()
    ref = np.array([orc.expectation_value(want, qs, m) for qs, m in expectation_cases(n)])
    for r in res:
        assert np.abs(r["evs"] - ref).max() < 2e-5, (r["evs"], ref)
        assert np.array_equal(r["evs"], res[0]["evs"])
    assert int(res[0]["swaps_after"]) > int(res[0]["swaps"])


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_peer_memory_swap_path(world, tmp_path):
    """the in-place peer-memory swap path (_swap_p2p): victims at any local bit, k > 1 exchanges, and the
    local SWAP pass that first lifts victims sitting on the lowest bits (MIN_SWAP_BIT)."""
    n = 10
    g = world.bit_length() - 1
    ops = random_ops(n, 48, seed=10 + world, max_local=n - g)
    want = oracle_full(n, ops)
    got, res = run_sharded(world, n, ops, tmp_path, p2p=True)
    assert np.abs(got - want).max() < 2e-6
    assert abs(float(res[0]["norm"]) - 1.0) < 1e-5
    assert int(res[0]["swaps"]) >= 2 and int(res[0]["swaps"]) == int(res[0]["nplan"])
    assert int(res[0]["local_swap_passes"]) >= 1  # random victims do land on the low bits
    check_expectations(n, want, res)


@pytest.mark.parametrize("world,mode", [(2, "remap"), (4, "remap"), (4, True)])
def test_library_schedule_replayed_over_gloo(world, mode, tmp_path):
    """The schedule of the library's planner (reordered gates + exchanges, qb200_sv_plan through the C ABI)
    replayed on oracle-driven shards over gloo with the index arithmetic of the push exchange (remap) and of
    the in-place exchange: the sharded result equals the in-order unsharded run."""
    from qsim_b200 import sv
    n = 10
    g = world.bit_length() - 1
    ops = random_ops(n, 60, seed=20 + world, max_local=n - g)
    want = oracle_full(n, ops)
    schedule = sv.plan(n, g, ops, reorder=True)
    assert sorted(s[1] for s in schedule if s[0] == "gate") == list(range(len(ops)))
    got, res = run_sharded(world, n, ops, tmp_path, p2p=mode, schedule=schedule)
    assert np.abs(got - want).max() < 3e-6
    assert int(res[0]["swaps"]) == sum(1 for s in schedule if s[0] == "swap") >= 1
    check_expectations(n, want, res)


def check_schedule(n, g, ops, schedule, glob=None):
    """every gate exactly once, per-qubit program order kept, targets local when the gate runs"""
    glob = set(range(n - g, n)) if glob is None else set(glob)
    last = {}
    seen = []
    for step in schedule:
        if step[0] == "swap":
            assert len(step[1]) == len(step[2]) >= 1
            assert set(step[2]) <= glob and not (set(step[1]) & glob)
            glob = (glob - set(step[2])) | set(step[1])
            continue
        i = step[1]
        seen.append(i)
        assert not (set(ops[i].qubits) & glob), "target on a global qubit"
        for q in list(ops[i].qubits) + list(ops[i].controls):
            assert last.get(q, -1) < i, "per-qubit order violated"
            last[q] = i
    assert sorted(seen) == list(range(len(ops)))
    return sum(1 - 2.0 ** -len(s[1]) for s in schedule if s[0] == "swap")


def test_library_planner_on_the_benchmark_circuits():
    from qsim_b200 import sv
    # (qubits, global qubits) -> exchanged bytes in shards, in-order and reordered: the reordering planner
    # needs 1-2 exchanges where the in-order one needs 3-4
    for n, g, max_cost in ((31, 1, 0.5), (32, 2, 0.75), (33, 3, 1.375), (36, 2, 0.75), (37, 3, 1.375)):
        _, ops = read_trace(os.path.join(ROOT, "tests", "golden", f"rqc_q{n}_d20_f4.trace"))
        c_in = check_schedule(n, g, ops, sv.plan(n, g, ops, reorder=False))
        c_re = check_schedule(n, g, ops, sv.plan(n, g, ops, reorder=True))
        assert c_re <= max_cost + 1e-9 and c_re < c_in
        # in-order schedules keep the program order
        assert [s[1] for s in sv.plan(n, g, ops, reorder=False) if s[0] == "gate"] == list(range(len(ops)))
    # nothing to do without global qubits; a given initial global set is honoured
    _, ops = read_trace(os.path.join(ROOT, "tests", "golden", "rqc_q20_d20_f4.trace"))
    assert all(s[0] == "gate" for s in sv.plan(20, 0, ops))
    check_schedule(20, 2, ops, sv.plan(20, 2, ops, global_qubits=[3, 7]), glob=[3, 7])
    # controls may stay global: a controlled gate whose control is global needs no exchange
    from qsim_b200.trace import TraceOp
    x = np.array([0, 0, 1, 0, 1, 0, 0, 0], np.float32)
    sched = sv.plan(6, 1, [TraceOp([0], [5], 1, x), TraceOp([1], [], 0, x)])
    assert all(s[0] == "gate" for s in sched)


def test_initial_global_set_is_chosen_for_the_circuit():
    """qb200_sv_plan_initial (csrc/sv_plan.h BestInitial): on a fresh state the qubit map is free, so the library picks
    the initial global qubits whose schedule exchanges the fewest shards -- never worse than the default (top g qubits),
    which it keeps on ties."""
    from qsim_b200 import sv
    from qsim_b200.trace import TraceOp
    x = np.array([0, 0, 1, 0, 1, 0, 0, 0], np.float32)

    def cost(steps):
        return sum(1 - 0.5 ** len(s[1]) for s in steps if s[0] == "swap")

    # every gate sits on the top qubits, qubits 0 and 1 are never touched: with them global no exchange is needed
    ops = [TraceOp([q], [], 0, x) for q in (7, 6, 5, 7, 4, 6, 3, 2)]
    for g in (1, 2):
        init = sv.plan_initial(8, g, ops)
        assert set(init) <= {0, 1} and len(init) == g
        assert cost(sv.plan(8, g, ops, global_qubits=init)) == 0 < cost(sv.plan(8, g, ops))
    # the benchmark circuits: the default set is already optimal within the search, and is kept
    for n, g in ((31, 1), (32, 2), (33, 3)):
        _, rops = read_trace(os.path.join(ROOT, "tests", "golden", f"rqc_q{n}_d20_f4.trace"))
        init = sv.plan_initial(n, g, rops)
        assert cost(sv.plan(n, g, rops, global_qubits=init)) <= cost(sv.plan(n, g, rops)) + 1e-12
        assert init == list(range(n - g, n))


def test_sharded_rqc_trace_world2(tmp_path):
    n, ops = read_trace(os.path.join(ROOT, "tests", "golden", "rqc_q20_d20_f4.trace"))
    ops = ops[:14]
    want = oracle_full(n, ops)
    got, res = run_sharded(2, n, ops, tmp_path)
    assert np.abs(got - want).max() < 2e-6
    # bytes per swap = shard * (1 - 2^-k) with k = 1 on two ranks
    assert int(res[0]["bytes_sent"]) == int(res[0]["swaps"]) * (8 << 19) // 2


def test_planner_properties():
    n, ops = read_trace(os.path.join(ROOT, "tests", "golden", "rqc_q37_d20_f4.trace"))
    opq = [list(o.qubits) + list(o.controls) for o in ops]
    for g in (1, 2, 3):
        plan = plan_swaps(opq, n, g)
        glob = set(range(n - g, n))
        it = iter(plan)
        step = next(it, None)
        for i, qs in enumerate(opq):
            if step is not None and step.before_op == i:
                assert set(step.incoming) <= glob and not (set(step.victims) & glob)
                assert not (set(step.victims) & set(qs))
                glob = (glob - set(step.incoming)) | set(step.victims)
                step = next(it, None)
            assert not (set(qs) & glob), "op touches a global qubit after planning"
        # far fewer swaps than gates that touch the initially-global qubits
        naive = sum(1 for qs in opq if set(qs) & set(range(n - g, n)))
        assert len(plan) <= naive
    assert plan_swaps(opq, n, 0) == []


def test_reindex_matrix():
    rs = np.random.RandomState(0)
    m = (rs.standard_normal((8, 8)) + 1j * rs.standard_normal((8, 8))).astype(np.complex64)
    from oracle.oracle import Oracle
    orc = Oracle()
    st = (rs.standard_normal(64) + 1j * rs.standard_normal(64)).astype(np.complex64)
    # gate on logical qubits (0,1,2) living at physical bits (4,1,3)
    bits, m2 = reindex_matrix(m, [4, 1, 3])
    assert bits == [1, 3, 4]
    a = st.copy(); orc.apply_gate(a, bits, m2)
    # reference: permute the state so that physical (4,1,3) become (0,1,2), apply, permute back
    perm = [4, 1, 3, 0, 2, 5]
    idx = np.arange(64)
    src = np.zeros(64, dtype=np.int64)
    for newbit, oldbit in enumerate(perm):
        src |= ((idx >> newbit) & 1) << oldbit
    b = st[src].copy(); orc.apply_gate(b, [0, 1, 2], m)
    back = np.empty_like(b); back[src] = b
    assert np.abs(a - back).max() < 1e-5
    # float (interleaved) input gives the same answer
    bits_f, m2f = reindex_matrix(m.reshape(-1).view(np.float32).copy(), [4, 1, 3])
    assert np.array_equal(m2f.view(np.complex64).reshape(8, 8), m2)


# ---- index arithmetic of the chunked exchange (csrc/sharded.cu), restated in numpy ----------------------------------
def _new_layout_reference(nl, lbits, my):
    """where amplitude i of a shard goes: (destination value of the victim bits, index in the destination buffer) --
    the definition (victim bits squeezed out, order of the rest kept, this shard's rank-bit value on top)"""
    k = len(lbits)
    i = np.arange(1 << nl, dtype=np.int64)
    v = np.zeros_like(i)
    packed = np.zeros_like(i)
    w = 0
    for b in range(nl):
        if b in lbits:
            v |= ((i >> b) & 1) << lbits.index(b)
        else:
            packed |= ((i >> b) & 1) << w
            w += 1
    return v, (my << (nl - k)) | packed


def _push_kernel_walk(nl, T, lbits, my, chunk_bits):
    """k_remap_push / k_remap_push_tma: every chunk's tile counters -> (source tile, destination, offset), with the
    chunk bits pinned through chunk_counter (run_overlapped computes cpos the same way)."""
    k = len(lbits)
    kl = sum(1 for b in lbits if b < T)
    kh = k - kl
    sub_bits = T - kl
    my_high = my >> kl
    tiles = 1 << (nl - T)
    c = len(chunk_bits)
    cpos = [kh + (cb - T) - sum(1 for b in lbits[kl:] if b < cb) for cb in chunk_bits]
    dst_v = np.full(1 << nl, -1, np.int64)
    dst_i = np.full(1 << nl, -1, np.int64)
    for cval in range(1 << c):
        for cc in range(tiles >> c):
            cnt = cc
            for j in range(c):   # chunk_counter
                lo = cnt & ((1 << cpos[j]) - 1)
                cnt = (((cnt >> cpos[j]) << 1 | ((cval >> j) & 1)) << cpos[j]) | lo
            v_high = (cnt & ((1 << kh) - 1)) ^ my_high
            tau = cnt >> kh
            for j in range(kl, k):
                b = lbits[j] - T
                lo = tau & ((1 << b) - 1)
                tau = (((tau >> b) << 1 | ((v_high >> (j - kl)) & 1)) << b) | lo
            packed_high = cnt >> kh
            for a in range(1 << T):   # amplitude a of the tile: sorted by its low victim bits
                v, r = 0, a
                for j in range(kl - 1, -1, -1):
                    b = lbits[j]
                    v |= ((r >> b) & 1) << j
                    r = ((r >> (b + 1)) << b) | (r & ((1 << b) - 1))
                src = (tau << T) | a
                assert dst_v[src] == -1
                dst_v[src] = v | (v_high << kl)
                dst_i[src] = (my << (nl - k)) | (packed_high << sub_bits) | r
                # the chunk bits of the source index carry cval
                for j, cb in enumerate(chunk_bits):
                    assert (src >> cb) & 1 == (cval >> j) & 1
    return dst_v, dst_i


def _copy_engine_boxes(nl, lbits, my, chunk_bits):
    """ce_push: free-bit runs between the pinned bits; the longest run above the row is the height of a 2-D copy, the
    others are looped over; the destination has the same runs at the packed positions."""
    k = len(lbits)
    dpos = lambda bit: bit - sum(1 for b in lbits if b < bit)
    dst_v = np.full(1 << nl, -1, np.int64)
    dst_i = np.full(1 << nl, -1, np.int64)
    sp = sorted(lbits + chunk_bits)
    runs, lo = [], 0
    for b in sp:
        if b > lo:
            runs.append((lo, b - lo))
        lo = b + 1
    if nl > lo:
        runs.append((lo, nl - lo))
    assert runs and runs[0][0] == 0
    boxed = max(range(1, len(runs)), key=lambda r: (runs[r][1], -r)) if len(runs) > 1 else None
    loops = [runs[r] for r in range(1, len(runs)) if r != boxed]
    width = 1 << runs[0][1]
    for cval in range(1 << len(chunk_bits)):
        src_fixed = sum(((cval >> j) & 1) << cb for j, cb in enumerate(chunk_bits))
        dst_fixed = (my << (nl - k)) | sum(((cval >> j) & 1) << dpos(cb) for j, cb in enumerate(chunk_bits))
        for v in range(1 << k):
            src_v = src_fixed | sum(((v >> j) & 1) << lbits[j] for j in range(k))
            nloop = 1 << sum(ln for _, ln in loops)
            for it in range(nloop):
                so, do, rest = src_v, dst_fixed, it
                for start, ln in loops:
                    val = rest & ((1 << ln) - 1)
                    rest >>= ln
                    so |= val << start
                    do |= val << dpos(start)
                height = 1 << runs[boxed][1] if boxed is not None else 1
                spitch = 1 << runs[boxed][0] if boxed is not None else 0
                dpitch = 1 << dpos(runs[boxed][0]) if boxed is not None else 0
                for h in range(height):
                    s0, d0 = so + h * spitch, do + h * dpitch
                    assert np.all(dst_v[s0:s0 + width] == -1)
                    dst_v[s0:s0 + width] = v
                    dst_i[s0:s0 + width] = np.arange(d0, d0 + width)
    return dst_v, dst_i


def test_chunked_exchange_index_arithmetic():
    """Both ways of moving a shard chunk by chunk -- the push kernels' tile walk with pinned counter bits and the copy
    engines' pitched boxes -- cover every amplitude exactly once and put it where the definition of the new layout
    says, for victims inside / above the tile, adjacent or apart, 0-3 chunk bits."""
    nl, T = 16, 6
    cases = [([15], [13, 14]), ([2], [14, 15]), ([3, 9], [12, 15]), ([7, 8, 12], [10, 14, 15]), ([0, 1], []),
             ([13, 14, 15], [11, 12]), ([5], [15])]
    for lbits, chunk_bits in cases:
        for my in (0, (1 << len(lbits)) - 1):
            want_v, want_i = _new_layout_reference(nl, lbits, my)
            got_v, got_i = _push_kernel_walk(nl, T, lbits, my, chunk_bits)
            assert np.array_equal(got_v, want_v) and np.array_equal(got_i, want_i), (lbits, chunk_bits)
    for lbits, chunk_bits in [([12], [14, 15]), ([12, 13], [15]), ([13, 15], [12, 14]), ([14], []), ([12, 14, 15], [13])]:
        for my in (0, (1 << len(lbits)) - 1):
            want_v, want_i = _new_layout_reference(nl, lbits, my)
            got_v, got_i = _copy_engine_boxes(nl, lbits, my, chunk_bits)
            assert np.array_equal(got_v, want_v) and np.array_equal(got_i, want_i), (lbits, chunk_bits)

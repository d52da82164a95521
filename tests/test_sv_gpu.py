"""Sharded state of the C ABI (qb200_sv_*, csrc/sharded.cu) on ONE GPU: 2, 4 and 8 shards on the same device --
the peers' "remote" memory is another allocation on the same GPU, so a one-GPU box runs the real exchange
kernels (k_remap_push, k_p2p_swap), both barrier kinds, the planner and every StateSpace / Simulator member
against the oracle.  Indexing and permutation results are compared bit-exactly."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from conftest import random_state, random_unitary  # noqa: E402


def make(shards, n, dtype=np.float32, **opts):
    from qsim_b200.sv import ShardedStateB200
    st = ShardedStateB200.single_process([0] * shards, n, dtype)
    for k, v in opts.items():
        st.set_option(k, v)
    return st


def raw_shards(st):
    """physical contents: [2^g, 2^n_local] complex array read straight from the shard buffers"""
    from qsim_b200 import _lib
    lib = _lib.load()
    nl = st.num_local_qubits()
    cd = np.complex64 if st.fp_type == np.float32 else np.complex128
    out = np.zeros((st.num_shards(), 1 << nl), dtype=cd)
    st.sync()
    for rank, dev, ptr, ctx in st.shards():
        buf = np.zeros(2 << nl, dtype=st.fp_type)
        assert lib.qb200_copy_d2h(ctx, st._dt, ptr, buf.ctypes.data_as(C.c_void_p), 2 << nl) == 0
        out[rank] = buf.view(cd)
    return out.reshape(-1)


def logical_from_physical(phys_state, pos):
    n = len(pos)
    idx = np.arange(1 << n, dtype=np.int64)
    p = np.zeros_like(idx)
    for q in range(n):
        p |= ((idx >> q) & 1) << int(pos[q])
    return phys_state[p]


@pytest.mark.parametrize("shards", [2, 4, 8])
@pytest.mark.parametrize("swap_mode", [1, 0, "tma"])
@pytest.mark.parametrize("barrier_flags", [0, 1])
def test_exchange_is_an_exact_index_permutation(shards, swap_mode, barrier_flags):
    """qb200_sv_swap for k = 1..g victims at low / high / mixed local bits: the logical state is unchanged bit
    for bit, and the physical buffers hold exactly the permutation the qubit map announces."""
    g = shards.bit_length() - 1
    n = 12 + g
    nl = n - g
    want = random_state(n, np.complex64, 5)
    opts = {"swap_mode": 1, "push_kernel": 1} if swap_mode == "tma" else {"swap_mode": swap_mode}  # bulk-copy push kernel
    swap_mode = 1 if swap_mode == "tma" else swap_mode
    st = make(shards, n, barrier_flags=barrier_flags, **opts)
    st.from_numpy(want)
    rng = np.random.default_rng(shards * 10 + swap_mode)
    cases = [[0], [nl - 1], [1, 2, 3][:g], [nl - 1, nl - 2, nl - 3][:g], [0, 5, nl - 1][:g], [3]]
    for victims_phys in cases:
        pos = st.qubit_map()
        at = {p: q for q, p in enumerate(pos)}
        victims = [at[p] for p in victims_phys]
        glob = [at[nl + t] for t in range(g)]
        incoming = list(rng.permutation(glob)[:len(victims)])
        st.Swap(victims, [int(q) for q in incoming])
        pos = st.qubit_map()
        assert sorted(pos) == list(range(n))
        for v in victims:
            assert pos[v] >= nl
        for q in incoming:
            assert pos[q] < nl
        got = logical_from_physical(raw_shards(st), pos)
        assert np.array_equal(got, want), f"victims at {victims_phys}"
    stats = st.stats()
    assert stats["swaps"] == len(cases)
    assert stats["bytes_sent_per_shard"] > 0
    # canonical order again: the buffers ARE the logical state
    st.Canonicalize()
    assert st.qubit_map() == list(range(n))
    assert np.array_equal(raw_shards(st), want)
    assert np.array_equal(st.to_numpy(), want)
    st.close()


@pytest.mark.parametrize("shards,dtype", [(2, np.float32), (4, np.float32), (8, np.float32), (4, np.float64)])
def test_copy_engine_exchange_is_the_same_permutation(shards, dtype):
    """push_kernel = 2 (csrc/sharded.cu ce_push): with the victims at bit 12 or above the exchange is a set of pitched
    2-D / 3-D copies on the copy engines -- bit for bit the permutation the qubit map announces, for victims that are
    adjacent, apart, at the top, and mixed with one below bit 12 (falls back to the push kernel)."""
    g = shards.bit_length() - 1
    nl = 17
    n = nl + g
    cd = np.complex64 if dtype == np.float32 else np.complex128
    want = random_state(n, cd, 9)
    st = make(shards, n, dtype, swap_mode=1, push_kernel=2)
    st.from_numpy(want)
    rng = np.random.default_rng(shards)
    cases = [[12], [nl - 1], [13, 14, 15][:g], [nl - 1, nl - 3, 12][:g], [14], [3, nl - 1][:g]]
    expect_ce = 0
    for victims_phys in cases:
        pos = st.qubit_map()
        at = {p: q for q, p in enumerate(pos)}
        victims = [at[p] for p in victims_phys]
        glob = [at[nl + t] for t in range(g)]
        incoming = [int(q) for q in rng.permutation(glob)[:len(victims)]]
        st.Swap(victims, incoming)
        expect_ce += min(victims_phys) >= 12
        pos = st.qubit_map()
        got = logical_from_physical(raw_shards(st), pos)
        assert np.array_equal(got, want), f"victims at {victims_phys}"
    stats = st.stats()
    assert stats["swaps"] == len(cases) and stats["copy_engine_swaps"] == expect_ce
    st.close()


def random_ops(n, count, seed, max_targets):
    from qsim_b200.trace import TraceOp
    rs = np.random.RandomState(seed)
    ops = []
    for i in range(count):
        g = int(rs.randint(1, max_targets + 1))
        qs = sorted(rs.choice(n, g, replace=False).tolist())
        u = random_unitary(g, seed * 1000 + i, np.complex64)
        cs, cv = [], 0
        if i % 4 == 3 and g <= 3:
            free = [q for q in range(n) if q not in qs]
            cs = sorted(rs.choice(free, int(rs.randint(1, 3)), replace=False).tolist())
            cv = int(rs.randint(0, 1 << len(cs)))
        ops.append(TraceOp(qs, cs, cv, np.ascontiguousarray(u).reshape(-1).view(np.float32).copy()))
    return ops


def oracle_run(oracle, n, ops, cdtype=np.complex64):
    st = np.zeros(1 << n, cdtype)
    st[0] = 1
    for op in ops:
        m = op.matrix.view(np.complex64).astype(cdtype)
        if op.controls:
            oracle.apply_controlled_gate(st, op.qubits, op.controls, op.cvals, m)
        else:
            oracle.apply_gate(st, op.qubits, m)
    return st


@pytest.mark.parametrize("shards,swap_mode,reorder", [(2, 1, 1), (4, 1, 1), (8, 1, 1), (4, 0, 1), (8, 0, 0), (2, 1, 0)])
def test_run_matches_unsharded_oracle(oracle, shards, swap_mode, reorder):
    """qb200_sv_run (planner + exchanges + single-GPU kernels on every shard) against the oracle on the
    unsharded state: gates of 1..6 qubits anywhere, controlled gates with local and global controls."""
    g = shards.bit_length() - 1
    n = 14 + g
    ops = random_ops(n, 70, seed=shards + 7 * swap_mode, max_targets=6)
    want = oracle_run(oracle, n, ops)
    st = make(shards, n, swap_mode=swap_mode, reorder=reorder)
    st.SetStateZero()
    st.Run(ops)
    got = st.to_numpy()
    assert np.abs(got - want).max() < 3e-6
    assert abs(st.Norm() - 1) < 1e-5
    stats = st.stats()
    assert stats["swaps"] >= 1 and stats["gate_passes"] == len(ops)
    # the online path (one ApplyGate at a time, least-recently-used victims) gives the same state
    st2 = make(shards, n, swap_mode=swap_mode)
    st2.SetStateZero()
    for op in ops:
        if op.controls:
            st2.ApplyControlledGate(op.qubits, op.controls, op.cvals, op.matrix)
        else:
            st2.ApplyGate(op.qubits, op.matrix)
    assert np.abs(st2.to_numpy() - want).max() < 3e-6
    assert abs(st.InnerProduct(st2) - 1) < 1e-5  # different qubit maps: both are re-mapped first
    st.close(); st2.close()


@pytest.mark.parametrize("shards,chunks_log2,dtype", [(2, 2, np.float32), (4, 1, np.float32), (4, 3, np.float32),
                                                      (8, 2, np.float32), (2, 2, np.float64), (4, 3, np.float64)])
def test_exchange_overlapped_with_the_last_gates_of_the_epoch(oracle, shards, chunks_log2, dtype):
    """qb200_sv_run, option overlap (csrc/sharded.cu run_overlapped): the last gate passes before an exchange run
    chunk by chunk -- the chunk bits as extra controls -- and every chunk is pushed on a second stream while the
    gates work on the next one.  Same schedule; a pass under extra controls may take another kernel of the
    dispatcher (different summation order), so the state equals the one of overlap = 0, and the oracle's, to
    round-off."""
    g = shards.bit_length() - 1
    n = 17 + g
    cd = np.complex64 if dtype == np.float32 else np.complex128
    ops = random_ops(n, 60, seed=31 + shards + chunks_log2, max_targets=5 if dtype == np.float32 else 4)
    want = oracle_run(oracle, n, ops, cd)
    states = []
    for overlap in (1, 0):
        st = make(shards, n, dtype, overlap=overlap, overlap_chunks_log2=chunks_log2, overlap_ce=shards != 8)
        st.SetStateZero()
        st.Run(ops)
        stats = st.stats()
        assert stats["swaps"] >= 1 and stats["gate_passes"] == len(ops)
        if overlap:
            assert stats["overlapped_swaps"] >= 1 and stats["overlapped_gate_passes"] >= stats["overlapped_swaps"]
        else:
            assert stats["overlapped_swaps"] == 0
        states.append(st.to_numpy())
        st.close()
    tol = 3e-6 if dtype == np.float32 else 1e-13
    assert np.abs(states[0] - want).max() < tol
    assert np.abs(states[0] - states[1]).max() < tol


def test_fresh_state_gets_the_qubit_map_the_circuit_wants(oracle):
    """|0...0> looks the same under every qubit map: qb200_sv_run relabels the map of a fresh state so that the
    circuit's best initial global qubits are global (no data moves).  Gates only on the top qubits: the default map
    needs exchanges, the chosen one none; the state is the oracle's either way, and a state that is no longer fresh
    keeps its map."""
    from qsim_b200.trace import TraceOp
    n, shards = 14, 4
    rs = np.random.RandomState(4)
    ops = []
    for i in range(12):
        qs = sorted(rs.choice(np.arange(4, n), 2, replace=False).tolist())   # qubits 0..3 are never touched
        ops.append(TraceOp(qs, [], 0, np.ascontiguousarray(random_unitary(2, i, np.complex64)).reshape(-1).view(np.float32).copy()))
    want = oracle_run(oracle, n, ops)
    st = make(shards, n)
    st.SetStateZero()
    st.Run(ops)
    assert st.stats()["swaps"] == 0 and sorted(q for q, p in enumerate(st.qubit_map()) if p >= n - 2) in ([0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3])
    assert np.abs(st.to_numpy() - want).max() < 3e-6
    pinned = make(shards, n, free_initial_map=0)
    pinned.SetStateZero()
    pinned.Run(ops)
    assert pinned.stats()["swaps"] >= 1
    assert np.abs(pinned.to_numpy() - want).max() < 3e-6
    # a state that has been touched is not fresh: its map stays, the same circuit needs its exchanges
    touched = make(shards, n)
    touched.SetStateZero()
    h = random_unitary(1, 99, np.complex64)
    touched.ApplyGate([5], h)
    touched.Run(ops)
    assert touched.stats()["swaps"] >= 1
    first = TraceOp([5], [], 0, np.ascontiguousarray(h).reshape(-1).view(np.float32).copy())
    assert np.abs(touched.to_numpy() - oracle_run(oracle, n, [first] + ops)).max() < 3e-6
    touched.close()
    st.close(); pinned.close()


@pytest.mark.parametrize("push_kernel", [0, 1])
def test_run_fp64(oracle, push_kernel):
    n, shards = 13, 4
    ops = random_ops(n, 40, seed=3, max_targets=5)
    want = oracle_run(oracle, n, ops, np.complex128)
    st = make(shards, n, np.float64, push_kernel=push_kernel)
    st.SetStateZero()
    st.Run(ops)
    assert np.abs(st.to_numpy() - want).max() < 1e-13
    assert st.stats()["swaps"] >= 1
    st.close()


@pytest.mark.parametrize("shards", [2, 8])
def test_statespace_members_on_a_sharded_state(oracle, shards):
    g = shards.bit_length() - 1
    n = 15 + g
    ops = random_ops(n, 30, seed=11, max_targets=4)
    want = oracle_run(oracle, n, ops)
    st = make(shards, n)
    st.SetStateZero()
    st.Run(ops)           # leaves a non-trivial qubit map behind
    assert st.qubit_map() != list(range(n))
    # GetAmpl / SetAmpl through the map
    for i in (0, 1, 5, (1 << n) - 1, 12345 % (1 << n)):
        assert abs(st.GetAmpl(i) - want[i]) < 3e-6
    # expectation values: local, global and mixed operator qubits
    for qs in ([0], [n - 1], [1, n - 2], [0, 3, n - 1], [2, 4, 5, 7]):
        m = random_unitary(len(qs), 99 + len(qs), np.complex64)
        assert abs(st.ExpectationValue(qs, m) - oracle.expectation_value(want, qs, m)) < 2e-5
    # norm, sampling: the sorted draws give the oracle's indices (cumulative sums in canonical order)
    norm = st.Norm()
    assert abs(norm - oracle.norm(want)) < 1e-5
    from qsim_b200 import StateSpaceB200
    rs = StateSpaceB200(np.float32).GenerateRandomValues(200, 3, norm)
    got_s = st.SampleWithValues(rs)
    want_s = oracle.sample(st.to_numpy(), rs)
    assert np.mean(got_s == want_s) > 0.97  # a draw within round-off of a cumulative sum may land next door
    assert np.array_equal(st.Sample(50, 7), st.Sample(50, 7))
    # measurement: PartialNorms / FindMeasuredBits / Collapse, the reference's lib/statespace.h:85-140 flow
    pn = st.PartialNorms()
    assert abs(pn.sum() - norm) < 1e-5
    cs = np.cumsum(pn)
    r = 0.37 * cs[-1]
    m = int(np.searchsorted(cs, r, side="left"))
    mask = (1 << 2) | (1 << (n - 1)) | (1 << 7)
    bits = st.FindMeasuredBits(m, r - (cs[m - 1] if m else 0.0), mask)
    full = st.to_numpy()
    csum = np.cumsum(np.abs(full.astype(np.complex128)) ** 2)
    k = int(np.searchsorted(csum, r, side="right"))
    assert bits == (k & mask) or bits == ((k + 1) & mask) or bits == ((k - 1) & mask)
    # scramble the map again, then collapse in that layout
    st.ApplyGate([n - 1], random_unitary(1, 5, np.complex64))
    want2 = st.to_numpy().copy()
    st.ApplyGate([n - 2, n - 1], np.eye(4, dtype=np.complex64))
    sel = (np.arange(1 << n) & mask) == bits
    pnorm = st.Collapse(mask, bits)
    exp = np.where(sel, want2, 0)
    assert abs(pnorm - np.sum(np.abs(exp.astype(np.complex128)) ** 2)) < 1e-5
    exp = (exp / np.sqrt(pnorm)).astype(np.complex64)
    assert np.abs(st.to_numpy() - exp).max() < 3e-6
    # SetStateUniform, BulkSetAmpl, Multiply, Add, Copy
    st.SetStateUniform()
    assert np.allclose(st.to_numpy(), 2.0 ** (-n / 2))
    st.ApplyGate([n - 1], np.eye(2, dtype=np.complex64))  # moves a qubit, map no longer canonical
    st.BulkSetAmpl(mask, bits, 1 + 2j)
    ref = np.full(1 << n, 2.0 ** (-n / 2), np.complex64)
    ref[sel] = 1 + 2j
    assert np.array_equal(st.to_numpy(), ref)
    st.BulkSetAmpl(mask, bits, -3j, exclude=True)
    ref[~sel] = -3j
    assert np.array_equal(st.to_numpy(), ref)
    st.Multiply(0.5)
    other = make(shards, n)
    other.CopyFrom(st)
    other.Add(st)
    assert np.array_equal(other.to_numpy(), ref)  # 0.5 ref + 0.5 ref
    st.close(); other.close()


def test_gate_too_large_for_a_shard_is_unsupported():
    from qsim_b200 import _lib
    st = make(8, 8)   # 5 local qubits
    st.SetStateZero()
    m = np.eye(64, dtype=np.complex64)
    q = (C.c_uint * 6)(0, 1, 2, 3, 4, 5)
    rc = _lib.load().qb200_sv_apply_gate(st._h, q, 6, m.view(np.float32).ctypes.data_as(C.c_void_p))
    assert rc == _lib.ERR_UNSUPPORTED
    assert st.GetAmpl(0) == 1
    st.close()
    with pytest.raises(Exception):
        make(8, 4)    # fewer than two local qubits (lib/multiprocess_custatevecex.h:160-163)


def test_state_on_another_device_is_refused():
    """qb200_state_alloc_on / the device check of gate passes (a context never touches another GPU's state)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import qsim_b200
    from qsim_b200 import _lib
    lib = _lib.load()
    sim1 = qsim_b200.SimulatorB200(np.float32, device=1)
    ss0 = qsim_b200.StateSpaceB200(np.float32, device=0)
    st0 = ss0.Create(10)
    q = (C.c_uint * 1)(0)
    m = np.eye(2, dtype=np.complex64)
    rc = lib.qb200_apply_gate(sim1._ctx, 0, st0.get(), 10, q, 1, m.view(np.float32).ctypes.data_as(C.c_void_p))
    assert rc == _lib.ERR_INVALID

"""Python mirror of the sharded state of the C ABI (qb200_sv_*, csrc/sharded.cu).

  ShardedStateB200.single_process(devices, n)   this process drives every shard
  ShardedStateB200.multi_process(dist, n)       one process per GPU; `dist` is an initialised torch.distributed
                                                (used only for the three host-side collectives of qb200_comm:
                                                handle exchange, sums of a few doubles, barriers)
Qubit arguments are logical qubits; the library owns the qubit map, the exchange kernels and the swap planner.
Reference precedent: the multi-device State + wire ordering of lib/vectorspace_custatevecex.h:163-287 and the
CuStateVecExRunner loop (lib/run_custatevecex.h:243-305).  No CPU fallback.
"""
import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import OK, QB200Error
from .backend import _dtype_code, _uarr


def pack_gates(ops, dtype=np.float32):
    """ops: objects with .qubits, .controls, .cvals, .matrix -> (ctypes array of qb200_gate, keep-alive list)."""
    arr = (_lib.Gate * max(len(ops), 1))()
    keep = []
    for i, op in enumerate(ops):
        q, nq = _uarr(op.qubits)
        c, nc = _uarr(getattr(op, "controls", ()) or ())
        m = np.ascontiguousarray(np.asarray(op.matrix).view(np.dtype(dtype)) if np.iscomplexobj(op.matrix)
                                 else np.asarray(op.matrix, dtype=dtype)).reshape(-1)
        if m.size != 2 << (2 * nq):
            raise ValueError("matrix size does not match the number of target qubits")
        keep += [q, c, m]
        arr[i].num_targets, arr[i].qs = nq, C.cast(q, _lib._pu)
        arr[i].num_controls, arr[i].cqs = nc, C.cast(c, _lib._pu)
        arr[i].cvals = int(getattr(op, "cvals", 0) or 0)
        arr[i].matrix = m.ctypes.data_as(C.c_void_p)
    return arr, keep


def plan(num_qubits: int, num_global: int, ops, global_qubits: Optional[Sequence[int]] = None, reorder: bool = True):
    """The library's swap planner alone (host only, no GPU): list of ("gate", op_index) /
    ("swap", victims, incoming)."""
    lib = _lib.load()
    arr, keep = pack_gates(ops)
    gq = None
    if global_qubits is not None:
        gq, _ = _uarr(global_qubits)
    need = C.c_uint64()
    rc = lib.qb200_sv_plan(num_qubits, num_global, arr, len(ops), gq, int(reorder), None, 0, C.byref(need))
    if rc != OK:
        raise QB200Error(rc, "qb200_sv_plan")
    buf = (C.c_int64 * max(need.value, 1))()
    rc = lib.qb200_sv_plan(num_qubits, num_global, arr, len(ops), gq, int(reorder), buf, need.value, C.byref(need))
    if rc != OK:
        raise QB200Error(rc, "qb200_sv_plan")
    out, w = [], 0
    while w < need.value:
        v = buf[w]; w += 1
        if v >= 0:
            out.append(("gate", int(v)))
        else:
            k = -v
            out.append(("swap", [int(buf[w + j]) for j in range(k)], [int(buf[w + k + j]) for j in range(k)]))
            w += 2 * k
    return out


def plan_initial(num_qubits: int, num_global: int, ops, reorder: bool = True):
    """The initial global qubits qb200_sv_run picks for this circuit on a fresh state (qb200_sv_plan_initial)."""
    lib = _lib.load()
    arr, keep = pack_gates(ops)
    out = (C.c_uint * max(num_global, 1))()
    rc = lib.qb200_sv_plan_initial(num_qubits, num_global, arr, len(ops), int(reorder), out)
    if rc != OK:
        raise QB200Error(rc, "qb200_sv_plan_initial")
    return [int(out[j]) for j in range(num_global)]


class _TorchComm:
    """qb200_comm over torch.distributed (host buffers; staged through a tensor on `device` for NCCL)."""

    def __init__(self, dist, device, group=None):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        world = dist.get_world_size()
        kw = {} if group is None else {"group": group}

        def allgather(_user, send, recv, nbytes):
            try:
                src = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_ubyte)), shape=(nbytes,))
                t = torch.from_numpy(src.copy()).to(device)
                outs = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(outs, t, **kw)
                dst = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_ubyte)), shape=(nbytes * world,))
                dst[:] = torch.cat(outs).cpu().numpy()
                return 0
            except Exception:  # never let an exception cross the C boundary
                return 1

        def allreduce(_user, buf, count):
            try:
                a = np.ctypeslib.as_array(buf, shape=(count,))
                t = torch.from_numpy(a.copy()).to(device)
                dist.all_reduce(t, **kw)
                a[:] = t.cpu().numpy()
                return 0
            except Exception:
                return 1

        def barrier(_user):
            try:
                dist.barrier(**kw)
                return 0
            except Exception:
                return 1

        self.struct = _lib.Comm(None, _lib.Comm.ALLGATHER(allgather), _lib.Comm.ALLREDUCE(allreduce),
                                _lib.Comm.BARRIER(barrier))


class ShardedStateB200:
    def __init__(self, handle, dtype, comm=None):
        self._lib = _lib.load()
        self._h = handle
        self._dt = _dtype_code(dtype)
        self.fp_type = np.dtype(dtype).type
        self._comm = comm  # keeps the callbacks alive

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def single_process(cls, devices: Sequence[int], num_qubits: int, dtype=np.float32):
        lib = _lib.load()
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = lib.qb200_sv_create(dev, len(devices), num_qubits, _dtype_code(dtype), C.byref(h))
        if rc != OK:
            raise QB200Error(rc, "qb200_sv_create", "(no CUDA device? the engine has no CPU fallback)")
        return cls(h, dtype)

    @classmethod
    def multi_process(cls, dist, num_qubits: int, device_index: int, dtype=np.float32, host_group=None):
        """host_group: a process group whose backend takes CPU tensors (gloo) for the host-side collectives --
        a few doubles each; without it they are staged through a device tensor of the default (NCCL) group."""
        import torch
        lib = _lib.load()
        if host_group is not None:
            comm = _TorchComm(dist, torch.device("cpu"), host_group)
        else:
            dev = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
            comm = _TorchComm(dist, dev)
        h = C.c_void_p()
        rc = lib.qb200_sv_create_mp(device_index, dist.get_rank(), dist.get_world_size(), C.byref(comm.struct),
                                    num_qubits, _dtype_code(dtype), C.byref(h))
        if rc != OK:
            raise QB200Error(rc, "qb200_sv_create_mp")
        return cls(h, dtype, comm)

    def close(self):
        if self._h:
            self._lib.qb200_sv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != OK:
            detail = ""
            if rc == _lib.ERR_CUDA:
                detail = f"cudaError {self._lib.qb200_sv_last_cuda_error(self._h)}"
            raise QB200Error(rc, what, detail)

    # ---- bookkeeping ------------------------------------------------------------------------------
    def num_qubits(self) -> int:
        return int(self._lib.qb200_sv_num_qubits(self._h))

    def num_shards(self) -> int:
        return int(self._lib.qb200_sv_num_shards(self._h))

    def num_local_qubits(self) -> int:
        return int(self._lib.qb200_sv_num_local_qubits(self._h))

    def qubit_map(self) -> List[int]:
        n = self.num_qubits()
        pos = (C.c_uint * n)()
        self._check(self._lib.qb200_sv_qubit_map(self._h, pos), "qubit_map")
        return list(pos)

    def shards(self):
        """[(rank, device, device pointer, ctx handle)] of the shards this process owns."""
        out = []
        for i in range(self._lib.qb200_sv_num_local_shards(self._h)):
            r, d, p, c = C.c_uint(), C.c_int(), C.c_void_p(), C.c_void_p()
            self._check(self._lib.qb200_sv_shard(self._h, i, C.byref(r), C.byref(d), C.byref(p), C.byref(c)), "shard")
            out.append((r.value, d.value, p.value, c.value))
        return out

    def set_option(self, key: str, value: int):
        self._check(self._lib.qb200_sv_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def sync(self):
        self._check(self._lib.qb200_sv_sync(self._h), "sync")

    def launch_count(self) -> int:
        return int(self._lib.qb200_sv_launch_count(self._h))

    def stats(self) -> dict:
        st = _lib.SvStats()
        self._check(self._lib.qb200_sv_get_stats(self._h, C.byref(st)), "get_stats")
        return {k: getattr(st, k) for k, _ in _lib.SvStats._fields_}

    def reset_stats(self):
        self._check(self._lib.qb200_sv_reset_stats(self._h), "reset_stats")

    # ---- StateSpace -------------------------------------------------------------------------------------
    def SetAllZeros(self):
        self._check(self._lib.qb200_sv_set_all_zeros(self._h), "SetAllZeros")

    def SetStateZero(self, reset_map: bool = False):
        if reset_map:
            self._check(self._lib.qb200_sv_reset_map(self._h), "reset_map")
        self._check(self._lib.qb200_sv_set_state_zero(self._h), "SetStateZero")

    def SetStateUniform(self):
        self._check(self._lib.qb200_sv_set_state_uniform(self._h), "SetStateUniform")

    def GetAmpl(self, i: int) -> complex:
        out = (C.c_double * 2)()
        self._check(self._lib.qb200_sv_get_ampl(self._h, i, out), "GetAmpl")
        return complex(out[0], out[1])

    def GetAmpls(self, indices) -> List[complex]:
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        out = np.zeros(2 * idx.size)
        self._check(self._lib.qb200_sv_get_ampls(self._h, idx.ctypes.data_as(_lib._pu64), idx.size,
                                                 out.ctypes.data_as(_lib._pd)), "GetAmpls")
        return [complex(out[2 * j], out[2 * j + 1]) for j in range(idx.size)]

    def SetAmpl(self, i: int, val: complex):
        self._check(self._lib.qb200_sv_set_ampl(self._h, i, complex(val).real, complex(val).imag), "SetAmpl")

    def BulkSetAmpl(self, mask: int, bits: int, val: complex, exclude: bool = False):
        self._check(self._lib.qb200_sv_bulk_set_ampl(self._h, mask, bits, complex(val).real, complex(val).imag,
                                                     int(exclude)), "BulkSetAmpl")

    def Norm(self) -> float:
        out = C.c_double()
        self._check(self._lib.qb200_sv_norm(self._h, C.byref(out)), "Norm")
        return out.value

    def InnerProduct(self, other: "ShardedStateB200") -> complex:
        out = (C.c_double * 2)()
        self._check(self._lib.qb200_sv_inner_product(self._h, other._h, out), "InnerProduct")
        return complex(out[0], out[1])

    def Add(self, src: "ShardedStateB200"):
        self._check(self._lib.qb200_sv_add(src._h, self._h), "Add")

    def CopyFrom(self, src: "ShardedStateB200"):
        self._check(self._lib.qb200_sv_copy(src._h, self._h), "Copy")

    def Multiply(self, a: float):
        self._check(self._lib.qb200_sv_multiply(self._h, float(a)), "Multiply")

    def SampleWithValues(self, sorted_rs) -> np.ndarray:
        rs = np.ascontiguousarray(sorted_rs, dtype=np.float64)
        out = np.zeros(rs.size, dtype=np.uint64)
        self._check(self._lib.qb200_sv_sample(self._h, rs.ctypes.data_as(_lib._pd), rs.size,
                                              out.ctypes.data_as(_lib._pu64)), "Sample")
        return out

    def Sample(self, num_samples: int, seed: int) -> np.ndarray:
        """lib/statespace_cuda.h:243-312: host RNG exactly as the reference draws it."""
        if num_samples == 0:
            return np.zeros(0, dtype=np.uint64)
        rs = np.empty(num_samples)
        self._lib.qb200_generate_random_values(num_samples, seed, self.Norm(), rs.ctypes.data_as(_lib._pd))
        return self.SampleWithValues(rs)

    def PartialNorms(self) -> np.ndarray:
        out = np.zeros(int(self._lib.qb200_sv_partial_norms_count(self._h)))
        self._check(self._lib.qb200_sv_partial_norms(self._h, out.ctypes.data_as(_lib._pd)), "PartialNorms")
        return out

    def FindMeasuredBits(self, m: int, r: float, mask: int) -> int:
        out = C.c_uint64()
        self._check(self._lib.qb200_sv_find_measured_bits(self._h, m, r, mask, C.byref(out)), "FindMeasuredBits")
        return int(out.value)

    def Collapse(self, mask: int, bits: int) -> float:
        out = C.c_double()
        self._check(self._lib.qb200_sv_collapse(self._h, mask, bits, C.byref(out)), "Collapse")
        return out.value

    def to_numpy(self) -> np.ndarray:
        """whole state as complex numbers, normal order (single-process states)."""
        n = self.num_qubits()
        host = np.zeros(2 << n, dtype=self.fp_type)
        self._check(self._lib.qb200_sv_copy_to_host(self._h, host.ctypes.data_as(C.c_void_p)), "copy_to_host")
        return host.view(np.complex64 if self.fp_type == np.float32 else np.complex128)

    def from_numpy(self, a):
        a = np.ascontiguousarray(a, dtype=np.complex64 if self.fp_type == np.float32 else np.complex128)
        if a.size != 1 << self.num_qubits():
            raise ValueError("state size mismatch")
        self._check(self._lib.qb200_sv_copy_from_host(self._h, a.ctypes.data_as(C.c_void_p)), "copy_from_host")

    # ---- Simulator ------------------------------------------------------------------------------------------
    def _matrix(self, matrix, nq):
        m = np.asarray(matrix)
        if np.iscomplexobj(m):
            m = np.ascontiguousarray(m, dtype=np.complex64 if self.fp_type == np.float32 else np.complex128).view(self.fp_type)
        m = np.ascontiguousarray(m, dtype=self.fp_type).reshape(-1)
        if m.size != 2 << (2 * nq):
            raise ValueError("matrix size does not match the number of target qubits")
        return m

    def ApplyGate(self, qs, matrix):
        q, nq = _uarr(qs)
        m = self._matrix(matrix, nq)
        rc = self._lib.qb200_sv_apply_gate(self._h, q, nq, m.ctypes.data_as(C.c_void_p))
        if rc != _lib.ERR_UNSUPPORTED:
            self._check(rc, "ApplyGate")

    def ApplyControlledGate(self, qs, cqs, cvals, matrix):
        q, nq = _uarr(qs)
        c, nc = _uarr(cqs)
        m = self._matrix(matrix, nq)
        rc = self._lib.qb200_sv_apply_controlled_gate(self._h, q, nq, c, nc, cvals, m.ctypes.data_as(C.c_void_p))
        if rc != _lib.ERR_UNSUPPORTED:
            self._check(rc, "ApplyControlledGate")

    def ExpectationValue(self, qs, matrix) -> complex:
        q, nq = _uarr(qs)
        m = self._matrix(matrix, nq)
        out = (C.c_double * 2)()
        rc = self._lib.qb200_sv_expectation_value(self._h, q, nq, m.ctypes.data_as(C.c_void_p), out)
        if rc != _lib.ERR_UNSUPPORTED:
            self._check(rc, "ExpectationValue")
        return complex(out[0], out[1])

    def Run(self, ops=None, packed=None):
        """a whole fused circuit (objects with .qubits/.controls/.cvals/.matrix, or pack_gates output)."""
        if packed is None:
            packed = pack_gates(ops, self.fp_type)
        arr, keep = packed
        count = len(keep) // 3
        self._check(self._lib.qb200_sv_run(self._h, arr, count), "Run")

    def Swap(self, victims, incoming):
        v, k = _uarr(victims)
        i, k2 = _uarr(incoming)
        assert k == k2
        self._check(self._lib.qb200_sv_swap(self._h, v, i, k), "Swap")

    def Canonicalize(self):
        self._check(self._lib.qb200_sv_canonicalize(self._h), "Canonicalize")

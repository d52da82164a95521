"""Python mirror of qsim's backend interface over the qsim_b200 C ABI.

Same names, argument meaning and error behaviour as the reference classes:
  State           <- VectorSpaceCUDA::Vector        lib/vectorspace_cuda.h:52-85
  StateSpaceB200  <- StateSpaceCUDA / StateSpace    lib/statespace_cuda.h:43-468, lib/statespace.h:32-141
  SimulatorB200   <- SimulatorCUDA                  lib/simulator_cuda.h:35-267
Host-side logic that the reference keeps on the host (RNG, prefix sums of the
partial norms) is restated here; all state-vector arithmetic runs in
libqsim_b200.so on the GPU.  There is no CPU fallback.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ERR_OOM, ERR_UNSUPPORTED, F32, F64, OK, QB200Error

_NP = {F32: np.float32, F64: np.float64}


def _dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise TypeError(f"unsupported fp_type {dtype}")


def _uarr(xs: Sequence[int]):
    xs = [int(x) for x in xs]
    return (C.c_uint * max(len(xs), 1))(*xs), len(xs)


class State:
    """Device state vector (VectorSpaceCUDA::Vector).  2*2^n scalars, normal order."""

    def __init__(self, ptr: Optional[int], num_qubits: int, dtype_code: int, owned: bool = True):
        self._ptr = ptr
        self._n = num_qubits
        self._dt = dtype_code
        self._owned = owned

    def get(self) -> Optional[int]:
        return self._ptr

    def num_qubits(self) -> int:
        return self._n

    def release(self) -> Optional[int]:
        p, self._ptr, self._n, self._owned = self._ptr, None, 0, False
        return p

    @staticmethod
    def requires_copy_to_host() -> bool:
        return True

    def __del__(self):
        if getattr(self, "_owned", False) and self._ptr:
            try:
                _lib.load().qb200_state_free(self._ptr)
            except Exception:
                pass
            self._ptr = None


@dataclass
class MeasurementResult:
    """lib/statespace.h:46-68"""
    mask: int = 0
    bits: int = 0
    bitstring: List[int] = field(default_factory=list)
    valid: bool = False


class _Base:
    def __init__(self, dtype=np.float32, device: int = -1):
        self._lib = _lib.load()
        self.fp_type = np.dtype(dtype).type
        self._dt = _dtype_code(dtype)
        ctx = C.c_void_p()
        rc = self._lib.qb200_ctx_create(device, C.byref(ctx))
        if rc != OK:
            raise QB200Error(rc, "qb200_ctx_create", "(no CUDA device? the engine has no CPU fallback)")
        self._ctx = ctx

    def __del__(self):
        ctx = getattr(self, "_ctx", None)
        if ctx:
            self._lib.qb200_ctx_destroy(ctx)
            self._ctx = None

    def _check(self, rc: int, what: str):
        if rc != OK:
            detail = ""
            if rc == _lib.ERR_CUDA:
                detail = self._lib.qb200_last_cuda_error_string(self._ctx).decode()
            raise QB200Error(rc, what, detail)

    def launch_count(self) -> int:
        return int(self._lib.qb200_launch_count(self._ctx))

    def ExpectationValuesSameQubits(self, qs, matrices, state: State) -> np.ndarray:
        """<psi|M_i|psi> for up to 8 operators on the same one or two qubits in one read pass
        (qb200_expectation_values_multi, csrc/expect_multi.cu)."""
        qs = [int(q) for q in qs]
        cdt = np.complex64 if self.fp_type == np.float32 else np.complex128
        ms = np.ascontiguousarray(np.stack([np.asarray(m, dtype=cdt).reshape(1 << len(qs), 1 << len(qs)) for m in matrices]))
        out = np.zeros(2 * len(matrices), dtype=np.float64)
        q = (C.c_uint * len(qs))(*qs)
        self._check(self._lib.qb200_expectation_values_multi(self._ctx, self._dt, state.get(), state.num_qubits(), q, len(qs),
                                                             ms.ctypes.data_as(C.c_void_p), len(matrices),
                                                             out.ctypes.data_as(C.POINTER(C.c_double))),
                    "ExpectationValuesSameQubits")
        return out.view(np.complex128)

    def last_kernel_name(self) -> str:
        """the gate / expectation kernel the dispatcher chose for this object's last pass"""
        return self._lib.qb200_last_kernel_name(self._ctx).decode()

    def set_tuning(self, key: str, value: int):
        self._check(self._lib.qb200_ctx_set_tuning(self._ctx, key.encode(), int(value)), "set_tuning")

    def set_stream(self, stream_handle: int):
        self._check(self._lib.qb200_ctx_set_stream(self._ctx, C.c_void_p(stream_handle)), "set_stream")

    def timer_start(self):
        self._check(self._lib.qb200_timer_start(self._ctx), "timer_start")

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.qb200_timer_stop_ms(self._ctx, C.byref(ms)), "timer_stop")
        return float(ms.value)


class StateSpaceB200(_Base):
    """Mirror of StateSpaceCUDA<FP> (lib/statespace_cuda.h) + StateSpace base (lib/statespace.h)."""

    # ---- VectorSpace ------------------------------------------------------
    @staticmethod
    def MinSize(num_qubits: int) -> int:
        return 2 << num_qubits

    def Create(self, num_qubits: int) -> State:
        p = C.c_void_p()
        rc = self._lib.qb200_state_alloc_on(self._ctx, num_qubits, self._dt, C.byref(p))  # on this object's device
        if rc == ERR_OOM:
            return self.Null()  # lib/vectorspace_cuda.h:90-95
        self._check(rc, "Create")
        return State(p.value, num_qubits, self._dt, owned=True)

    def CreateFromPointer(self, ptr: int, num_qubits: int) -> State:
        """Create(fp_type* p, n): wraps caller-owned device memory (lib/vectorspace_cuda.h:100-102)."""
        return State(ptr, num_qubits, self._dt, owned=False)

    def Null(self) -> State:
        return State(None, 0, self._dt, owned=False)

    @staticmethod
    def IsNull(state: State) -> bool:
        return state.get() is None

    def Copy(self, src, dest) -> bool:
        """Copy(state,state) / Copy(state,host) / Copy(host,state) (lib/vectorspace_cuda.h:112-160)."""
        if isinstance(src, State) and isinstance(dest, State):
            if src.num_qubits() != dest.num_qubits():
                return False
            self._check(self._lib.qb200_copy_d2d(self._ctx, self._dt, src.get(), dest.get(),
                                                 self.MinSize(src.num_qubits())), "Copy")
            return True
        if isinstance(src, State):
            count = self.MinSize(src.num_qubits())
            assert dest.dtype == _NP[self._dt] and dest.size >= count and dest.flags.c_contiguous
            self._check(self._lib.qb200_copy_d2h(self._ctx, self._dt, src.get(),
                                                 dest.ctypes.data_as(C.c_void_p), count), "Copy")
            return True
        src = np.ascontiguousarray(src, dtype=_NP[self._dt])
        count = min(src.size, self.MinSize(dest.num_qubits()))
        self._check(self._lib.qb200_copy_h2d(self._ctx, self._dt, src.ctypes.data_as(C.c_void_p),
                                             dest.get(), count), "Copy")
        return True

    def DeviceSync(self):
        self._check(self._lib.qb200_sync(self._ctx), "DeviceSync")

    # ---- convenience (not in the reference): whole state as complex numpy ----
    def to_numpy(self, state: State) -> np.ndarray:
        buf = np.empty(self.MinSize(state.num_qubits()), dtype=_NP[self._dt])
        self.Copy(state, buf)
        return buf.view(np.complex64 if self._dt == F32 else np.complex128)

    def from_numpy(self, amplitudes: np.ndarray, state: State):
        cdt = np.complex64 if self._dt == F32 else np.complex128
        a = np.ascontiguousarray(amplitudes, dtype=cdt)
        assert a.size == 1 << state.num_qubits()
        self.Copy(a.view(_NP[self._dt]), state)

    # ---- StateSpace -------------------------------------------------------
    def InternalToNormalOrder(self, state: State):
        self._check(self._lib.qb200_internal_to_normal_order(self._ctx, self._dt, state.get(), state.num_qubits()), "InternalToNormalOrder")

    def NormalToInternalOrder(self, state: State):
        self._check(self._lib.qb200_normal_to_internal_order(self._ctx, self._dt, state.get(), state.num_qubits()), "NormalToInternalOrder")

    def SetAllZeros(self, state: State):
        self._check(self._lib.qb200_set_all_zeros(self._ctx, self._dt, state.get(), state.num_qubits()), "SetAllZeros")

    def SetStateUniform(self, state: State):
        self._check(self._lib.qb200_set_state_uniform(self._ctx, self._dt, state.get(), state.num_qubits()), "SetStateUniform")

    def SetStateZero(self, state: State):
        self._check(self._lib.qb200_set_state_zero(self._ctx, self._dt, state.get(), state.num_qubits()), "SetStateZero")

    def GetAmpl(self, state: State, i: int) -> complex:
        out = (C.c_double * 2)()
        self._check(self._lib.qb200_get_ampl(self._ctx, self._dt, state.get(), i, out), "GetAmpl")
        return complex(out[0], out[1])

    def SetAmpl(self, state: State, i: int, re, im=None):
        if im is None:
            re, im = complex(re).real, complex(re).imag
        self._check(self._lib.qb200_set_ampl(self._ctx, self._dt, state.get(), i, re, im), "SetAmpl")

    def BulkSetAmpl(self, state: State, mask: int, bits: int, re, im=None, exclude: bool = False):
        if im is None:
            re, im = complex(re).real, complex(re).imag
        self._check(self._lib.qb200_bulk_set_ampl(self._ctx, self._dt, state.get(), state.num_qubits(),
                                                  mask, bits, re, im, int(bool(exclude))), "BulkSetAmpl")

    def Add(self, src: State, dest: State) -> bool:
        if src.num_qubits() != dest.num_qubits():
            return False
        self._check(self._lib.qb200_add(self._ctx, self._dt, src.get(), dest.get(), src.num_qubits()), "Add")
        return True

    def Multiply(self, a: float, state: State):
        self._check(self._lib.qb200_multiply(self._ctx, self._dt, float(a), state.get(), state.num_qubits()), "Multiply")

    def InnerProduct(self, state1: State, state2: State) -> complex:
        if state1.num_qubits() != state2.num_qubits():
            return complex(math.nan, 0.0)  # lib/statespace_cuda.h:221-223
        out = (C.c_double * 2)()
        self._check(self._lib.qb200_inner_product(self._ctx, self._dt, state1.get(), state2.get(),
                                                  state1.num_qubits(), out), "InnerProduct")
        return complex(out[0], out[1])

    def RealInnerProduct(self, state1: State, state2: State) -> float:
        if state1.num_qubits() != state2.num_qubits():
            return math.nan
        out = C.c_double()
        self._check(self._lib.qb200_real_inner_product(self._ctx, self._dt, state1.get(), state2.get(),
                                                       state1.num_qubits(), C.byref(out)), "RealInnerProduct")
        return out.value

    def Norm(self, state: State) -> float:
        out = C.c_double()
        self._check(self._lib.qb200_norm(self._ctx, self._dt, state.get(), state.num_qubits(), C.byref(out)), "Norm")
        return out.value

    def GenerateRandomValues(self, num_samples: int, seed: int, max_value: float) -> np.ndarray:
        """lib/util.h:67-85 (std::mt19937 + uniform_real_distribution, sorted)."""
        rs = np.empty(num_samples, dtype=np.float64)
        self._check(self._lib.qb200_generate_random_values(num_samples, seed, max_value,
                                                           rs.ctypes.data_as(C.POINTER(C.c_double))), "GenerateRandomValues")
        return rs

    def GenerateRandomValuesOnDevice(self, num_samples: int, seed: int, max_value: float) -> np.ndarray:
        """The same sorted values drawn by the device Mersenne Twister (csrc/sample_rng.cu), copied back."""
        rs = np.empty(num_samples, dtype=np.float64)
        self._check(self._lib.qb200_generate_random_values_device(self._ctx, num_samples, seed, max_value,
                                                                  rs.ctypes.data_as(C.POINTER(C.c_double))),
                    "GenerateRandomValuesOnDevice")
        return rs

    def Sample(self, state: State, num_samples: int, seed: int, host_rng: bool = False,
               norm: Optional[float] = None) -> np.ndarray:
        """lib/statespace_cuda.h:243-312: norm -> sorted random values -> device search.  The values are drawn on
        the device (the reference's TODO at :292; bit-identical to the host draw for the same norm,
        csrc/sample_rng.cu) unless host_rng asks for the reference's host path.  norm: upper bound of the draws;
        default = Norm(state) on the host path, the total of the sampler's own chunk sums on the device path."""
        out = np.zeros(num_samples, dtype=np.uint64)
        if num_samples > 0:
            if host_rng:
                norm = self.Norm(state) if norm is None else norm
                rs = self.GenerateRandomValues(num_samples, seed, norm)
                self.SampleWithValues(state, rs, out)
            else:
                norm = -1.0 if norm is None else norm
                self._check(self._lib.qb200_sample_seeded(self._ctx, self._dt, state.get(), state.num_qubits(),
                                                          num_samples, seed, norm,
                                                          out.ctypes.data_as(C.POINTER(C.c_uint64))), "Sample")
        return out

    def SampleWithValues(self, state: State, sorted_rs: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        rs = np.ascontiguousarray(sorted_rs, dtype=np.float64)
        if out is None:
            out = np.zeros(rs.size, dtype=np.uint64)
        self._check(self._lib.qb200_sample(self._ctx, self._dt, state.get(), state.num_qubits(),
                                           rs.ctypes.data_as(C.POINTER(C.c_double)), rs.size,
                                           out.ctypes.data_as(C.POINTER(C.c_uint64))), "Sample")
        return out

    def PartialNorms(self, state: State) -> np.ndarray:
        cnt = int(self._lib.qb200_partial_norms_count(state.num_qubits()))
        out = np.empty(cnt, dtype=np.float64)
        self._check(self._lib.qb200_partial_norms(self._ctx, self._dt, state.get(), state.num_qubits(),
                                                  out.ctypes.data_as(C.POINTER(C.c_double))), "PartialNorms")
        return out

    def FindMeasuredBits(self, m: int, r: float, mask: int, state: State) -> int:
        out = C.c_uint64()
        self._check(self._lib.qb200_find_measured_bits(self._ctx, self._dt, state.get(), state.num_qubits(),
                                                       m, r, mask, C.byref(out)), "FindMeasuredBits")
        return int(out.value)

    def Collapse(self, mr: MeasurementResult, state: State):
        self._check(self._lib.qb200_collapse(self._ctx, self._dt, state.get(), state.num_qubits(),
                                             mr.mask, mr.bits, None), "Collapse")

    def VirtualMeasure(self, qubits: Sequence[int], r01: float, state: State) -> MeasurementResult:
        """lib/statespace.h:97-140.  `r01` is a uniform [0,1) draw; the reference
        draws RandomValue(rgen, norm) = norm * u for the same u."""
        result = MeasurementResult(valid=True)
        for q in qubits:
            if q >= state.num_qubits():
                result.valid = False
                return result
            result.mask |= 1 << q
        csum = np.cumsum(self.PartialNorms(state))
        r = r01 * csum[-1]
        m = 0
        while r > csum[m]:
            m += 1
        if m > 0:
            r -= csum[m - 1]
        result.bits = self.FindMeasuredBits(m, r, result.mask, state)
        result.bitstring = [(result.bits >> q) & 1 for q in qubits]
        return result

    def Measure(self, qubits: Sequence[int], r01: float, state: State) -> MeasurementResult:
        """lib/statespace.h:85-95"""
        result = self.VirtualMeasure(qubits, r01, state)
        if result.valid:
            self.Collapse(result, state)
        return result


class SimulatorB200(_Base):
    """Mirror of SimulatorCUDA<FP> (lib/simulator_cuda.h:35-267)."""

    @staticmethod
    def SIMDRegisterSize() -> int:
        return 32  # lib/simulator_cuda.h:265-267 (tests derive their qubit ranges from it)

    def _matrix(self, matrix, num_targets):
        m = np.ascontiguousarray(matrix)
        if np.iscomplexobj(m):
            m = m.astype(np.complex64 if self._dt == F32 else np.complex128).reshape(-1).view(_NP[self._dt])
        else:
            m = m.astype(_NP[self._dt]).reshape(-1)
        if num_targets <= 6:
            assert m.size == 2 << (2 * num_targets), "matrix must be 2^G x 2^G complex"
        return m

    def ApplyGate(self, qs: Sequence[int], matrix, state: State):
        """lib/simulator_cuda.h:70-125.  Gates on more than 6 qubits are ignored like the reference."""
        q, nq = _uarr(qs)
        m = self._matrix(matrix, nq)
        rc = self._lib.qb200_apply_gate(self._ctx, self._dt, state.get(), state.num_qubits(), q, nq,
                                        m.ctypes.data_as(C.c_void_p))
        if rc != ERR_UNSUPPORTED:
            self._check(rc, "ApplyGate")

    def ApplyControlledGate(self, qs: Sequence[int], cqs: Sequence[int], cvals: int, matrix, state: State):
        """lib/simulator_cuda.h:135-207."""
        q, nq = _uarr(qs)
        c, nc = _uarr(cqs)
        m = self._matrix(matrix, nq)
        rc = self._lib.qb200_apply_controlled_gate(self._ctx, self._dt, state.get(), state.num_qubits(),
                                                   q, nq, c, nc, cvals, m.ctypes.data_as(C.c_void_p))
        if rc != ERR_UNSUPPORTED:
            self._check(rc, "ApplyControlledGate")

    def ExpectationValue(self, qs: Sequence[int], matrix, state: State) -> complex:
        """lib/simulator_cuda.h:216-260."""
        q, nq = _uarr(qs)
        m = self._matrix(matrix, nq)
        out = (C.c_double * 2)()
        rc = self._lib.qb200_expectation_value(self._ctx, self._dt, state.get(), state.num_qubits(), q, nq,
                                               m.ctypes.data_as(C.c_void_p), out)
        if rc != ERR_UNSUPPORTED:
            self._check(rc, "ExpectationValue")
        return complex(out[0], out[1])

    def OneQubitMoments(self, state: State) -> np.ndarray:
        """qb200_one_qubit_moments (csrc/moments.cu; no reference counterpart): array [n, 4] of
        S00, S11, Re S01, Im S01 per qubit -- every single-qubit reduced density matrix from 3-4 read-only
        passes.  <M_q> = m00 S00 + m11 S11 + m01 S01 + m10 conj(S01)."""
        n = state.num_qubits()
        out = np.zeros((max(n, 1), 4), dtype=np.float64)
        rc = self._lib.qb200_one_qubit_moments(self._ctx, self._dt, state.get(), n, out.ctypes.data_as(_lib._pd))
        self._check(rc, "OneQubitMoments")
        return out[:n]

    @staticmethod
    def moment_expectation(moments_q, matrix) -> complex:
        """<M> of a 2x2 operator from one row of OneQubitMoments."""
        m = np.asarray(matrix, dtype=np.complex128).reshape(2, 2)
        s01 = complex(moments_q[2], moments_q[3])
        return complex(m[0, 0] * moments_q[0] + m[1, 1] * moments_q[1] + m[0, 1] * s01 + m[1, 0] * s01.conjugate())

    def ExpectationValues(self, terms, state: State) -> List[complex]:
        """Batched expectation values (include/qsim_b200/expect_b200.h; no reference counterpart):
        `terms` = [(qs, matrix), ...]; every read-only pass is enqueued back to back and the values are
        read after ONE stream synchronisation.  Same kernels, same values as ExpectationValue per term."""
        terms = list(terms)
        self._check(self._lib.qb200_reduce_batch_begin(self._ctx, len(terms)), "reduce_batch_begin")
        try:
            for qs, matrix in terms:
                self.ExpectationValue(qs, matrix, state)
        finally:
            out = (C.c_double * (2 * len(terms) + 2))()
            count = C.c_uint32()
            rc = self._lib.qb200_reduce_batch_end(self._ctx, out, len(terms), C.byref(count))
        self._check(rc, "reduce_batch_end")
        return [complex(out[2 * i], out[2 * i + 1]) for i in range(count.value)]

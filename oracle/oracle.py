"""ctypes front ends of the CPU checkers (TEST INFRASTRUCTURE ONLY).

  Oracle   -- oracle/libqsim_oracle.so, our plain-C restatement of the
              reference's SimulatorBasic/StateSpaceBasic (qsim_oracle.c).
  RefEngine -- oracle/_ref/libqsim_ref_*.so, the UNMODIFIED reference CPU
              backends compiled in place from /root/reference (ref_shim.cc).
              Present only where `make -C oracle` ran with the reference tree
              available; the prebuilt files travel to the GPU box.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp, _u, _u64, _d = C.c_void_p, C.c_uint, C.c_uint64, C.c_double
_pu, _pd, _pu64 = C.POINTER(C.c_uint), C.POINTER(C.c_double), C.POINTER(C.c_uint64)


def build(with_ref: bool = True):
    """Compiles the checker (never the product).  `make` skips the reference
    part by itself when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", _HERE, "all" if with_ref else "libqsim_oracle.so"], check=True)


def _uarr(xs):
    xs = [int(x) for x in xs]
    return (C.c_uint * max(len(xs), 1))(*xs), len(xs)


def _suffix(state: np.ndarray) -> str:
    if state.dtype == np.complex64:
        return "f32"
    if state.dtype == np.complex128:
        return "f64"
    raise TypeError("oracle states are complex64/complex128 numpy arrays")


def _real_dtype(state):
    return np.float32 if state.dtype == np.complex64 else np.float64


def _nq(state: np.ndarray) -> int:
    n = int(state.size).bit_length() - 1
    assert state.size == 1 << n and state.flags.c_contiguous
    return n


class Oracle:
    """Plain-C restatement; operates in place on complex numpy arrays."""

    def __init__(self):
        path = os.path.join(_HERE, "libqsim_oracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        self.lib = C.CDLL(path)
        for suf in ("f32", "f64"):
            getattr(self.lib, f"orc_norm_{suf}").restype = _d
            getattr(self.lib, f"orc_real_inner_product_{suf}").restype = _d
            getattr(self.lib, f"orc_collapse_{suf}").restype = _d
            getattr(self.lib, f"orc_sample_norm_{suf}").restype = _d
            getattr(self.lib, f"orc_find_measured_bits_{suf}").restype = _u64

    def _m(self, matrix, state):
        cdt = state.dtype
        m = np.ascontiguousarray(matrix)
        if np.iscomplexobj(m):
            m = m.astype(cdt).reshape(-1).view(_real_dtype(state))
        else:
            m = m.astype(_real_dtype(state)).reshape(-1)
        return m

    def apply_gate(self, state, qs, matrix):
        return self.apply_controlled_gate(state, qs, [], 0, matrix)

    def apply_controlled_gate(self, state, qs, cqs, cvals, matrix):
        q, nq = _uarr(qs)
        c, nc = _uarr(cqs)
        m = self._m(matrix, state)
        fn = getattr(self.lib, f"orc_apply_controlled_gate_{_suffix(state)}")
        rc = fn(state.ctypes.data_as(_vp), _u(_nq(state)), q, _u(nq), c, _u(nc), _u64(cvals),
                m.ctypes.data_as(_vp))
        assert rc == 0
        return state

    def expectation_value(self, state, qs, matrix) -> complex:
        q, nq = _uarr(qs)
        m = self._m(matrix, state)
        out = (_d * 2)()
        fn = getattr(self.lib, f"orc_expectation_value_{_suffix(state)}")
        rc = fn(state.ctypes.data_as(_vp), _u(_nq(state)), q, _u(nq), m.ctypes.data_as(_vp), out)
        assert rc == 0
        return complex(out[0], out[1])

    def _call(self, name, state, *args):
        return getattr(self.lib, f"orc_{name}_{_suffix(state)}")(state.ctypes.data_as(_vp), _u(_nq(state)), *args)

    def set_state_zero(self, state):
        self._call("set_state_zero", state)

    def set_state_uniform(self, state):
        self._call("set_state_uniform", state)

    def norm(self, state) -> float:
        return self._call("norm", state)

    def inner_product(self, s1, s2) -> complex:
        out = (_d * 2)()
        getattr(self.lib, f"orc_inner_product_{_suffix(s1)}")(s1.ctypes.data_as(_vp), s2.ctypes.data_as(_vp), _u(_nq(s1)), out)
        return complex(out[0], out[1])

    def multiply(self, a, state):
        rt = C.c_float if state.dtype == np.complex64 else C.c_double
        getattr(self.lib, f"orc_multiply_{_suffix(state)}")(rt(a), state.ctypes.data_as(_vp), _u(_nq(state)))

    def add(self, src, dest):
        getattr(self.lib, f"orc_add_{_suffix(src)}")(src.ctypes.data_as(_vp), dest.ctypes.data_as(_vp), _u(_nq(src)))

    def bulk_set_ampl(self, state, mask, bits, val, exclude=False):
        rt = C.c_float if state.dtype == np.complex64 else C.c_double
        self._call("bulk_set_ampl", state, _u64(mask), _u64(bits), rt(complex(val).real), rt(complex(val).imag), C.c_int(int(exclude)))

    def sample_norm(self, state) -> float:
        return self._call("sample_norm", state)

    def sample(self, state, sorted_rs) -> np.ndarray:
        rs = np.ascontiguousarray(sorted_rs, dtype=np.float64)
        out = np.zeros(rs.size, dtype=np.uint64)
        self._call("sample", state, rs.ctypes.data_as(_pd), _u64(rs.size), out.ctypes.data_as(_pu64))
        return out

    def collapse(self, state, mask, bits) -> float:
        return self._call("collapse", state, _u64(mask), _u64(bits))

    def find_measured_bits(self, state, r, mask) -> int:
        return int(self._call("find_measured_bits", state, _d(r), _u64(mask)))


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_library_path():
    """Best prebuilt reference library this host CPU can execute, or None."""
    flags = _cpu_flags()
    for name, need in (("avx512", {"avx512f", "avx2", "fma", "bmi2"}), ("avx2", {"avx2", "fma"}), ("sse", {"sse4_1"})):
        p = os.path.join(_HERE, "_ref", f"libqsim_ref_{name}.so")
        if need <= flags and os.path.exists(p):
            return p
    return None


BASIC_F32, BASIC_F64, SIMD_F32 = 0, 1, 2


class RefEngine:
    """One state inside the unmodified reference CPU backend (see ref_shim.cc)."""
    _libs = {}

    def __init__(self, kind: int, num_qubits: int, threads: int = 1, path: str = None):
        path = path or ref_library_path()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libqsim_ref_*.so not built (needs the reference tree)")
        lib = RefEngine._libs.get(path)
        if lib is None:
            lib = C.CDLL(path)
            lib.ref_create.restype = _vp
            lib.ref_norm.restype = _d
            lib.ref_real_inner_product.restype = _d
            lib.ref_simd_name.restype = C.c_char_p
            RefEngine._libs[path] = lib
        self.lib = lib
        self.kind = kind
        self.n = num_qubits
        self.cdtype = np.complex128 if kind == BASIC_F64 else np.complex64
        self.h = lib.ref_create(kind, _u(num_qubits), _u(threads))
        if not self.h:
            raise MemoryError("reference state allocation failed")

    def simd_name(self) -> str:
        return self.lib.ref_simd_name().decode()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_destroy(_vp(self.h))
            self.h = None

    def _m(self, matrix):
        rdt = np.float64 if self.kind == BASIC_F64 else np.float32
        m = np.ascontiguousarray(matrix)
        if np.iscomplexobj(m):
            m = m.astype(self.cdtype).reshape(-1).view(rdt)
        else:
            m = m.astype(rdt).reshape(-1)
        return m

    def set_zero(self):
        self.lib.ref_set_zero(_vp(self.h))

    def set_uniform(self):
        self.lib.ref_set_uniform(_vp(self.h))

    def from_numpy(self, amps):
        a = np.ascontiguousarray(amps, dtype=self.cdtype)
        assert a.size == 1 << self.n
        self.lib.ref_from_normal(_vp(self.h), a.ctypes.data_as(_vp))

    def to_numpy(self):
        out = np.empty(1 << self.n, dtype=self.cdtype)
        self.lib.ref_to_normal(_vp(self.h), out.ctypes.data_as(_vp))
        return out

    def apply_gate(self, qs, matrix):
        q, nq = _uarr(qs)
        m = self._m(matrix)
        self.lib.ref_apply_gate(_vp(self.h), q, _u(nq), m.ctypes.data_as(_vp))

    def apply_controlled_gate(self, qs, cqs, cvals, matrix):
        q, nq = _uarr(qs)
        c, nc = _uarr(cqs)
        m = self._m(matrix)
        self.lib.ref_apply_controlled_gate(_vp(self.h), q, _u(nq), c, _u(nc), _u64(cvals), m.ctypes.data_as(_vp))

    def expectation_value(self, qs, matrix) -> complex:
        q, nq = _uarr(qs)
        m = self._m(matrix)
        out = (_d * 2)()
        self.lib.ref_expectation_value(_vp(self.h), q, _u(nq), m.ctypes.data_as(_vp), out)
        return complex(out[0], out[1])

    def norm(self) -> float:
        return self.lib.ref_norm(_vp(self.h))

    def inner_product(self, other) -> complex:
        out = (_d * 2)()
        self.lib.ref_inner_product(_vp(self.h), _vp(other.h), out)
        return complex(out[0], out[1])

    def real_inner_product(self, other) -> float:
        return self.lib.ref_real_inner_product(_vp(self.h), _vp(other.h))

    def multiply(self, a):
        self.lib.ref_multiply(_vp(self.h), _d(a))

    def add_from(self, src) -> bool:
        return self.lib.ref_add(_vp(src.h), _vp(self.h)) == 0

    def sample(self, num, seed) -> np.ndarray:
        out = np.zeros(num, dtype=np.uint64)
        self.lib.ref_sample(_vp(self.h), _u64(num), _u(seed), out.ctypes.data_as(_pu64))
        return out

    def measure(self, qs, seed):
        q, nq = _uarr(qs)
        mask, bits = _u64(), _u64()
        rc = self.lib.ref_measure(_vp(self.h), q, _u(nq), _u(seed), C.byref(mask), C.byref(bits))
        return (rc == 0, int(mask.value), int(bits.value))

    def collapse(self, mask, bits):
        self.lib.ref_collapse(_vp(self.h), _u64(mask), _u64(bits))

    def get_ampl(self, i) -> complex:
        out = (_d * 2)()
        self.lib.ref_get_ampl(_vp(self.h), _u64(i), out)
        return complex(out[0], out[1])

    def set_ampl(self, i, val):
        self.lib.ref_set_ampl(_vp(self.h), _u64(i), _d(complex(val).real), _d(complex(val).imag))

    def bulk_set_ampl(self, mask, bits, val, exclude=False):
        self.lib.ref_bulk_set_ampl(_vp(self.h), _u64(mask), _u64(bits), _d(complex(val).real), _d(complex(val).imag), C.c_int(int(exclude)))

    def generate_random_values(self, num, seed, max_value) -> np.ndarray:
        out = np.empty(num, dtype=np.float64)
        self.lib.ref_generate_random_values(_u64(num), _u(seed), _d(max_value), out.ctypes.data_as(_pd))
        return out

"""CPU restatement of csrc/sample_rng.cu (the device Mersenne Twister behind qb200_sample_seeded): the three-phase
parallel twist and libstdc++'s uniform_real_distribution<double> arithmetic, in numpy, against (a) numpy's MT19937
and (b) the library's host helper qb200_generate_random_values = std::mt19937 + std::uniform_real_distribution +
std::sort (lib/util.h:67-85).  The GPU test (tests/test_statespace_gpu.py) checks the kernel itself bit for bit."""
import ctypes as C

import numpy as np

N, M = 624, 397


def seed_state(seed):
    st = np.zeros(N, dtype=np.uint64)
    st[0] = seed & 0xffffffff
    for i in range(1, N):
        st[i] = (1812433253 * (int(st[i - 1]) ^ (int(st[i - 1]) >> 30)) + i) & 0xffffffff
    return st.astype(np.uint32)


def twist_word(cur, nxt, far):
    y = (cur & np.uint32(0x80000000)) | (nxt & np.uint32(0x7fffffff))
    return far ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908b0df), np.uint32(0))


def phased_twist(o):
    """the kernel's three phases: every word of a phase depends on OLD words and on NEW words of earlier phases only"""
    nw = np.zeros_like(o)
    d = N - M   # 227
    nw[:d] = twist_word(o[:d], o[1:d + 1], o[M:M + d])                       # phase 1: old words only
    nw[d:2 * d] = twist_word(o[d:2 * d], o[d + 1:2 * d + 1], nw[:d])         # phase 2: far = new words of phase 1
    i = np.arange(2 * d, N - 1)
    nw[i] = twist_word(o[i], o[i + 1], nw[i - d])                            # phase 3: far = new words of phase 2
    nw[N - 1] = twist_word(o[N - 1:N], nw[0:1], nw[N - 1 - d:N - d])[0]      # the last word wraps to NEW word 0
    return nw


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9d2c5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xefc60000))
    return y ^ (y >> np.uint32(18))


def device_algorithm(num, seed, max_value):
    st = seed_state(seed)
    out = []
    while len(out) < num:
        st = phased_twist(st)
        w = temper(st).astype(np.float64)
        s = w[0::2] + w[1::2] * 4294967296.0          # low word first; one rounding, like libstdc++'s generate_canonical
        c = s * 5.421010862427522170037e-20            # / 2^64, exact
        c = np.where(c >= 1.0, np.nextafter(1.0, 0.0), c)
        out.extend((c * max_value + 0.0).tolist())
    return np.sort(np.array(out[:num]))


def test_phased_twist_is_mt19937():
    for seed in (0, 1, 5489, 2 ** 32 - 1):
        mt = np.random.MT19937()
        state = mt.state
        state["state"]["key"], state["state"]["pos"] = seed_state(seed), N
        mt.state = state
        want = mt.random_raw(3 * N).astype(np.uint32)
        st = seed_state(seed)
        got = []
        for _ in range(3):
            st = phased_twist(st)
            got.append(temper(st))
        assert np.array_equal(np.concatenate(got), want)


def test_device_algorithm_equals_the_reference_host_sequence():
    from qsim_b200 import _lib
    lib = _lib.load()
    for seed, num, mx in ((1, 5, 1.0), (7, 311, 0.73), (7, 312, 1.0), (7, 313, 3.0), (123456, 5000, 0.99999), (2 ** 32 - 1, 1000, 1.0000001)):
        host = np.empty(num, dtype=np.float64)
        assert lib.qb200_generate_random_values(num, seed, mx, host.ctypes.data_as(C.POINTER(C.c_double))) == 0
        dev = device_algorithm(num, seed, mx)
        assert np.array_equal(host.view(np.uint64), dev.view(np.uint64)), (seed, num, mx)

"""Parity of the CUDA StateSpace path against the CPU oracle; mirrors
tests/statespace_testfixture.h (TestAdd :358, TestNormSmall :393,
TestNormAndInnerProduct :466, TestSamplingSmall :514, TestOrdering :612,
TestMeasurementSmall :731, TestCollapse :840, TestBulkSetAmplitude* :920-1044)."""
import numpy as np
import pytest

from conftest import random_state

pytestmark = pytest.mark.gpu
RDT = {np.complex64: np.float32, np.complex128: np.float64}
DTYPES = [np.complex64, np.complex128]


def make(cdt):
    import qsim_b200
    return qsim_b200.StateSpaceB200(RDT[cdt])


@pytest.mark.parametrize("cdt", DTYPES)
@pytest.mark.parametrize("n", [0, 1, 2, 5, 9, 14, 17])
def test_init_and_ampl_access(cdt, n):
    ss = make(cdt)
    st = ss.Create(n)
    assert not ss.IsNull(st) and st.num_qubits() == n
    assert ss.MinSize(n) == 2 << n
    ss.SetStateZero(st)
    a = ss.to_numpy(st)
    assert a[0] == 1 and np.count_nonzero(a) == 1
    ss.SetStateUniform(st)
    a = ss.to_numpy(st)
    v = RDT[cdt](1.0 / np.sqrt(float(1 << n)))
    assert np.all(a.real == v) and np.all(a.imag == 0)
    ss.SetAllZeros(st)
    assert np.count_nonzero(ss.to_numpy(st)) == 0
    # TestOrdering: amplitude i really is element i
    for i in {0, (1 << n) - 1, (1 << n) // 3}:
        ss.SetAmpl(st, i, complex(i + 1, -(i + 2)))
        assert ss.GetAmpl(st, i) == complex(i + 1, -(i + 2))
    ss.InternalToNormalOrder(st)
    a = ss.to_numpy(st)
    assert a[(1 << n) - 1] == complex(1 << n, -((1 << n) + 1))


@pytest.mark.parametrize("cdt", DTYPES)
@pytest.mark.parametrize("n", [0, 1, 4, 10, 15, 20])
def test_reductions_match_oracle(oracle, cdt, n):
    ss = make(cdt)
    h1, h2 = random_state(n, cdt, 1), random_state(n, cdt, 2)
    s1, s2 = ss.Create(n), ss.Create(n)
    ss.from_numpy(h1, s1)
    ss.from_numpy(h2, s2)
    tol = 1e-6 if cdt == np.complex64 else 1e-13
    assert abs(ss.Norm(s1) - oracle.norm(h1)) < tol
    assert abs(ss.InnerProduct(s1, s2) - oracle.inner_product(h1, h2)) < tol
    assert abs(ss.RealInnerProduct(s1, s2) - oracle.inner_product(h1, h2).real) < tol
    # deterministic reductions
    assert ss.Norm(s1) == ss.Norm(s1)
    # mismatch conventions (lib/statespace_cuda.h:221-223, 190-192)
    if n > 0:
        s3 = ss.Create(n - 1)
        assert np.isnan(ss.InnerProduct(s1, s3).real) and np.isnan(ss.RealInnerProduct(s1, s3))
        assert ss.Add(s1, s3) is False and ss.Copy(s1, s3) is False


@pytest.mark.parametrize("cdt", DTYPES)
@pytest.mark.parametrize("n", [0, 1, 7, 16])
def test_add_multiply_bulk_set(oracle, cdt, n):
    ss = make(cdt)
    h1, h2 = random_state(n, cdt, 3), random_state(n, cdt, 4)
    s1, s2 = ss.Create(n), ss.Create(n)
    ss.from_numpy(h1, s1)
    ss.from_numpy(h2, s2)
    assert ss.Add(s1, s2)
    w = h2.copy(); oracle.add(h1, w)
    assert np.array_equal(ss.to_numpy(s2), w)
    ss.Multiply(0.37, s2)
    oracle.multiply(0.37, w)
    assert np.array_equal(ss.to_numpy(s2), w)
    for mask, bits, excl in [(0, 0, False), (1, 1, False), (1, 0, True), (0b101 & ((1 << n) - 1), 0b100 & ((1 << n) - 1), False),
                             (((1 << n) - 1), ((1 << n) - 1) // 2, True)]:
        ss.BulkSetAmpl(s2, mask, bits, complex(0.5, -0.25), exclude=excl)
        oracle.bulk_set_ampl(w, mask, bits, complex(0.5, -0.25), excl)
        assert np.array_equal(ss.to_numpy(s2), w), (mask, bits, excl)
    d = ss.Create(n)
    assert ss.Copy(s2, d)
    assert np.array_equal(ss.to_numpy(d), w)


@pytest.mark.parametrize("cdt", DTYPES)
@pytest.mark.parametrize("n", [1, 3, 9, 13, 14, 18])
def test_collapse_matches_oracle(oracle, cdt, n):
    ss = make(cdt)
    import qsim_b200
    h = random_state(n, cdt, 5)
    st = ss.Create(n)
    ss.from_numpy(h, st)
    mask = 0b1011 & ((1 << n) - 1)
    bits = 0b0010 & mask
    ss.Collapse(qsim_b200.MeasurementResult(mask=mask, bits=bits, valid=True), st)
    w = h.copy()
    oracle.collapse(w, mask, bits)
    tol = 2e-7 if cdt == np.complex64 else 1e-14
    got = ss.to_numpy(st)
    assert np.abs(got - w).max() < tol
    assert np.array_equal(got == 0, w == 0)
    assert abs(ss.Norm(st) - 1) < 1e-6


@pytest.mark.parametrize("cdt", DTYPES)
@pytest.mark.parametrize("n", [1, 4, 12, 13, 15, 19])
def test_sample_matches_oracle(oracle, cdt, n):
    """same sorted random values -> same indices as the serial CPU scan, except
    where a value sits within rounding of a cumulative-sum boundary."""
    ss = make(cdt)
    h = random_state(n, cdt, 6)
    if n >= 4:
        h[3:11] = 0  # zero-probability amplitudes must never be returned
    st = ss.Create(n)
    ss.from_numpy(h, st)
    num = 20000
    norm = ss.Norm(st)
    rs = ss.GenerateRandomValues(num, 11, norm)
    assert np.all(np.diff(rs) >= 0) and rs[-1] < norm
    got = ss.SampleWithValues(st, rs)
    want = oracle.sample(h, rs)
    diff = np.nonzero(got != want)[0]
    if diff.size:
        csum = np.cumsum(np.abs(h.astype(np.complex128)) ** 2)
        for i in diff:
            lo, hi = sorted((int(got[i]), int(want[i])))
            assert hi - lo <= 1 or np.all(np.abs(h[lo + 1:hi]) == 0)
            assert abs(csum[lo] - rs[i]) < 1e-9
    assert diff.size <= 2
    assert np.all(np.abs(h[got.astype(np.int64)]) > 0)
    # API path with seed (Sample = Norm + GenerateRandomValues + search)
    assert np.array_equal(ss.Sample(st, num, 11, norm=norm), got)
    assert np.count_nonzero(ss.Sample(st, num, 11) != got) <= 2   # upper bound from the sampler's own chunk sums
    # tail: values beyond the total probability map to 2^n - 1 (lib/statespace_basic.h:227-229)
    tail = ss.SampleWithValues(st, np.array([norm * 0.5, norm * 1.5, norm * 2.0]))
    assert tail[1] == (1 << n) - 1 and tail[2] == (1 << n) - 1
    assert ss.Sample(st, 0, 1).size == 0


def test_generate_random_values_is_reference_sequence():
    """std::mt19937(1) + uniform_real_distribution<double>(0,1): first draws are fixed by the C++ standard."""
    ss = make(np.complex64)
    rs = ss.GenerateRandomValues(5, 1, 1.0)
    import random
    # independent restatement: generate_canonical<double,53> consumes two 32-bit words per draw
    class MT(random.Random):
        pass
    import numpy.random as npr
    mt = npr.MT19937()
    st = mt.state
    key = np.zeros(624, dtype=np.uint32)
    key[0] = 1
    for i in range(1, 624):
        key[i] = (1812433253 * (int(key[i - 1]) ^ (int(key[i - 1]) >> 30)) + i) & 0xffffffff
    st["state"]["key"], st["state"]["pos"] = key, 624
    mt.state = st
    words = mt.random_raw(10).astype(np.float64)
    want = np.sort([(words[2 * i] + words[2 * i + 1] * 4294967296.0) / 18446744073709551616.0 for i in range(5)])
    assert np.allclose(rs, want, rtol=0, atol=1e-16)


@pytest.mark.parametrize("seed", [0, 1, 12345, 2 ** 32 - 1])
def test_device_random_values_are_the_host_sequence_bit_for_bit(seed):
    """csrc/sample_rng.cu: device MT19937 + libstdc++'s uniform_real_distribution<double> arithmetic + radix sort
    against GenerateRandomValues<double> (lib/util.h:67-85, std::mt19937 on the host): identical doubles for counts
    around the 312-draws-per-twist boundary, for many twists, and for max_value != 1."""
    ss = make(np.complex64)
    for num, mx in [(1, 1.0), (5, 1.0), (311, 0.73), (312, 1.0), (313, 3.0), (624, 1.0 - 2.0 ** -30), (1000, 0.999999),
                    (100003, 1.0000001)]:
        host = ss.GenerateRandomValues(num, seed, mx)
        dev = ss.GenerateRandomValuesOnDevice(num, seed, mx)
        assert np.array_equal(host.view(np.uint64), dev.view(np.uint64)), (num, mx)


@pytest.mark.parametrize("cdt", DTYPES)
def test_sample_with_device_rng_equals_the_host_rng_path(oracle, cdt):
    """Sample(state, m, seed) draws its values on the device (the reference's TODO, lib/statespace_cuda.h:292):
    same indices as the host-RNG path and as the oracle's serial scan over the host values."""
    ss = make(cdt)
    n = 15
    h = random_state(n, cdt, 21)
    st = ss.Create(n)
    ss.from_numpy(h, st)
    norm = ss.Norm(st)
    for num, seed in [(1, 3), (777, 0), (50000, 99)]:
        dev = ss.Sample(st, num, seed, norm=norm)
        host = ss.Sample(st, num, seed, host_rng=True, norm=norm)
        assert np.array_equal(dev, host)            # same upper bound: same doubles, same indices
        want = oracle.sample(h, ss.GenerateRandomValues(num, seed, norm))
        assert np.count_nonzero(dev != want) <= 2   # a value within round-off of a cumulative sum may land next door
        # default: the upper bound is the total of the sampler's own chunk sums (no separate Norm pass) -- Norm(state)
        # up to the summation order, so at most a boundary case differs
        assert np.count_nonzero(ss.Sample(st, num, seed) != dev) <= 2


@pytest.mark.parametrize("cdt", DTYPES)
@pytest.mark.parametrize("n", [1, 5, 13, 16])
def test_measure_matches_oracle(oracle, cdt, n):
    """VirtualMeasure/Measure (lib/statespace.h:85-140) vs serial CPU scan + collapse."""
    ss = make(cdt)
    h = random_state(n, cdt, 7)
    tol = 3e-7 if cdt == np.complex64 else 1e-14
    for u in (0.0, 0.123, 0.5, 0.87, 0.999):
        st = ss.Create(n)
        ss.from_numpy(h, st)
        qubits = sorted({0, n // 2, n - 1})
        pn = ss.PartialNorms(st)
        assert pn.size == max(1, (1 << n) >> 13)
        assert abs(pn.sum() - 1) < 1e-6
        res = ss.Measure(qubits, u, st)
        assert res.valid
        mask = sum(1 << q for q in qubits)
        assert res.mask == mask
        want_bits = oracle.find_measured_bits(h, u * pn.sum(), mask)
        assert res.bits == want_bits
        assert res.bitstring == [(want_bits >> q) & 1 for q in qubits]
        w = h.copy()
        oracle.collapse(w, mask, want_bits)
        assert np.abs(ss.to_numpy(st) - w).max() < tol
    bad = ss.Measure([n], 0.5, st)
    assert not bad.valid

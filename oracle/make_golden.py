"""Generates tests/golden/kat_reference.npz from the UNMODIFIED reference CPU
backends (oracle/_ref/libqsim_ref_*.so, built by `make -C oracle` from
/root/reference).  Run from the repo root in the build container:

    python -m oracle.make_golden

Inputs are regenerated from seeds (numpy RandomState: frozen stream), outputs
are the reference's.  The committed file pins the C oracle (tests/test_oracle.py)
and, through it, the CUDA path, on machines where /root/reference is absent.
"""
import os

import numpy as np

from .oracle import BASIC_F32, BASIC_F64, SIMD_F32, RefEngine

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "kat_reference.npz")


def kat_state(n, cdtype, seed):
    rs = np.random.RandomState(seed)
    st = rs.standard_normal(1 << n) + 1j * rs.standard_normal(1 << n)
    st /= np.linalg.norm(st)
    return st.astype(cdtype)


def kat_matrix(g, seed, cdtype):
    rs = np.random.RandomState(1000 + seed)
    a = rs.standard_normal((1 << g, 1 << g)) + 1j * rs.standard_normal((1 << g, 1 << g))
    return (a / (1 << g)).astype(cdtype)


def kat_cases():
    """(name, n, targets, controls, cvals)"""
    cases = []
    for n in (1, 2, 4, 7, 10):
        for g in range(0, min(n, 6) + 1):
            rs = np.random.RandomState(n * 10 + g)
            layouts = {tuple(range(g)), tuple(range(n - g, n)), tuple(sorted(rs.choice(n, g, replace=False).tolist()))}
            for qs in sorted(layouts):
                cases.append((n, list(qs), [], 0))
    for n in (4, 7, 10):
        rs = np.random.RandomState(n)
        for g in range(0, 5):
            for c in (1, 2, 3):
                if g + c > n:
                    continue
                perm = rs.permutation(n)[: g + c].tolist()
                cvals = int(rs.randint(0, 1 << c))
                cases.append((n, sorted(perm[:g]), sorted(perm[g:]), cvals))
    return cases


def main():
    out = {}
    kinds = (("f32", BASIC_F32, np.complex64), ("f64", BASIC_F64, np.complex128), ("simd", SIMD_F32, np.complex64))
    for tag, kind, cdt in kinds:
        for idx, (n, qs, cqs, cvals) in enumerate(kat_cases()):
            st = kat_state(n, cdt, idx)
            m = kat_matrix(len(qs), idx, cdt)
            e = RefEngine(kind, n, 1)
            e.from_numpy(st)
            if len(qs) >= 1 and not cqs:
                ev = e.expectation_value(qs, m)
                out[f"{tag}/ev/{idx}"] = np.array([ev.real, ev.imag])
            if cqs:
                e.apply_controlled_gate(qs, cqs, cvals, m)
            else:
                e.apply_gate(qs, m)
            out[f"{tag}/gate/{idx}"] = e.to_numpy()
        # state-space known answers
        for idx, n in enumerate((1, 3, 8, 12)):
            a, b = kat_state(n, cdt, 500 + idx), kat_state(n, cdt, 600 + idx)
            ea, eb = RefEngine(kind, n, 1), RefEngine(kind, n, 1)
            ea.from_numpy(a)
            eb.from_numpy(b)
            ip = ea.inner_product(eb)
            out[f"{tag}/ss/{idx}/norm_ip"] = np.array([ea.norm(), ip.real, ip.imag, ea.real_inner_product(eb)])
            out[f"{tag}/ss/{idx}/samples"] = ea.sample(256, 7)
            ok, mask, bits = ea.measure(sorted({0, n - 1}), 3)
            out[f"{tag}/ss/{idx}/measure"] = np.array([int(ok), mask, bits], dtype=np.uint64)
            out[f"{tag}/ss/{idx}/collapsed"] = ea.to_numpy()
            eb.multiply(0.625)
            eb.add_from(ea)
            eb.bulk_set_ampl(1, 1, 0.25 - 0.5j, False)
            out[f"{tag}/ss/{idx}/mul_add_bulk"] = eb.to_numpy()
    e = RefEngine(BASIC_F32, 1, 1)
    out["rng/seed1_norm1"] = e.generate_random_values(64, 1, 1.0)
    out["rng/seed7_norm0.9"] = e.generate_random_values(64, 7, 0.9)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Random-quantum-circuit generator in qsim's text format, following the layout rules
the reference's circuits/circuit_q30 obeys (6-column grid, sites row-major, 8 CZ
layer patterns, H first, then T first / a random one of {x_1_2, y_1_2, t} differing from the
previous one on qubits leaving a CZ) and
extending them to any number of sites (partial last row), so the weak-scaling and
34/36/37-qubit workloads are built exactly like BASELINE config 2.

  python tools/gen_rqc.py N DEPTH [seed] > circuit_qN

The output is only an INPUT to the reference's parser + fuser (oracle/_ref/ref_fuse),
which writes the fused-gate traces committed under tests/golden/.
"""
import random
import sys

COLS = 6


def cz_layer(t, n):
    """CZ pairs of cycle t (1-based); patterns repeat with period 8."""
    k = (t - 1) % 8
    pairs = []
    if k in (0, 1, 4, 5):  # horizontal
        off = {0: 0, 1: 2, 4: 1, 5: 3}[k]
        a = off
        while a + 1 < n:
            if a // COLS == (a + 1) // COLS:
                pairs.append((a, a + 1))
            a += 4
    else:  # vertical
        rows = (n + COLS - 1) // COLS
        for r in range(rows - 1):
            if k in (2, 3) and r % 2 == 1:
                par = ((r - 1) // 2) % 2
                want = par if k == 2 else 1 - par
            elif k in (6, 7) and r % 2 == 0:
                par = (r // 2) % 2
                want = par if k == 6 else 1 - par
            else:
                continue
            for c in range(COLS):
                a, b = r * COLS + c, (r + 1) * COLS + c
                if c % 2 == want and b < n:
                    pairs.append((a, b))
    return pairs


def generate(n, depth, seed=0):
    rng = random.Random(seed)
    lines = [str(n)]
    for q in range(n):
        lines.append(f"0 h {q}")
    in_cz_prev = set()
    last_1q = {}
    for t in range(1, depth + 1):
        pairs = cz_layer(t, n)
        busy = set()
        for a, b in pairs:
            lines.append(f"{t} cz {a} {b}")
            busy.update((a, b))
        for q in sorted(in_cz_prev - busy):
            if q not in last_1q:
                g = "t"
            else:
                g = rng.choice([x for x in ("x_1_2", "y_1_2", "t") if x != last_1q[q]])
            last_1q[q] = g
            lines.append(f"{t} {g} {q}")
        in_cz_prev = busy
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    n, depth = int(sys.argv[1]), int(sys.argv[2])
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else n
    sys.stdout.write(generate(n, depth, seed))

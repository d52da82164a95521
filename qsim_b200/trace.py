"""Fused-gate trace files (written by oracle/ref_fuse.cc from the reference's own
parser + MultiQubitGateFuser): the exact ApplyGate/ApplyControlledGate call
sequence lib/run_qsim.h:264-280 issues for a circuit."""
import struct
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class TraceOp:
    qubits: List[int]
    controls: List[int]
    cvals: int
    matrix: np.ndarray  # float32, 2 * 4^len(qubits), row-major interleaved (re, im)


def read_trace(path: str):
    """Returns (num_qubits, [TraceOp])."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"QB2TRACE":
        raise ValueError(f"{path}: not a fused-gate trace")
    version, num_qubits, num_ops, fp_bytes = struct.unpack_from("<IIII", data, 8)
    if version != 1 or fp_bytes != 4:
        raise ValueError(f"{path}: unsupported trace version/precision")
    off = 24
    ops = []
    for _ in range(num_ops):
        nq, nc, cvals = struct.unpack_from("<IIQ", data, off)
        off += 16
        qs = list(struct.unpack_from(f"<{nq}I", data, off)); off += 4 * nq
        cs = list(struct.unpack_from(f"<{nc}I", data, off)); off += 4 * nc
        size = 2 << (2 * nq)
        m = np.frombuffer(data, dtype="<f4", count=size, offset=off).copy()
        off += 4 * size
        ops.append(TraceOp(qs, cs, cvals, m))
    return num_qubits, ops

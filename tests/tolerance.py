"""The parity tolerance of BASELINE.json's north star, written down once.

"values within 1e-12 relative (1e-14 absolute) in fp64".  An entry passes if
    |got - ref| <= 1e-12 * |ref|                       (relative), or
    |got - ref| <= 1e-14 * max(1, scale)               (absolute),
where `scale` is the largest |ref| in the same Jacobian row (for g: in the same
constraint set).  The scale matters because force-scaled rows hold entries up
to ~1e4, where one fp64 ulp (1.8e-12) already exceeds a bare 1e-14: an entry
that cancels to ~1e-2 inside such a row cannot be reproduced to 1e-14 by ANY
re-association, only to 1e-14 of the operands it was computed from.
`strict_fraction` reports how many entries miss the bare (scale = 1) criterion.
"""
import numpy as np

REL = 1e-12
ABS = 1e-14


def check_rows(got, ref, row_ptr, what="jac"):
    """got/ref: (B, nnz); row_ptr: CSR row pointer. Returns (n_bad, strict_fraction, worst)."""
    d = np.abs(got - ref)
    a = np.abs(ref)
    lens = np.diff(row_ptr)
    nz = lens > 0
    scale_rows = np.ones((ref.shape[0], len(lens)))
    if ref.shape[1]:
        scale_rows[:, nz] = np.maximum.reduceat(a, row_ptr[:-1][nz], axis=1)
    scale = np.repeat(np.maximum(scale_rows, 1.0), lens, axis=1)
    ok = (d <= REL * a) | (d <= ABS * scale)
    strict = (d <= REL * a) | (d <= ABS)
    worst = float((d / np.maximum(REL * a, ABS * scale)).max()) if d.size else 0.0
    return int((~ok).sum()), float((~strict).mean()) if d.size else 0.0, worst


def check_sets(got, ref, sets, what="g"):
    """got/ref: (B, m); sets: [(name, start, count)]."""
    d = np.abs(got - ref)
    a = np.abs(ref)
    scale = np.ones_like(a)
    for _, s, k in sets:
        if k:
            scale[:, s:s + k] = np.maximum(a[:, s:s + k].max(axis=1, keepdims=True), 1.0)
    ok = (d <= REL * a) | (d <= ABS * scale)
    strict = (d <= REL * a) | (d <= ABS)
    worst = float((d / np.maximum(REL * a, ABS * scale)).max()) if d.size else 0.0
    return int((~ok).sum()), float((~strict).mean()) if d.size else 0.0, worst

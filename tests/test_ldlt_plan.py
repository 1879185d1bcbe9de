"""CPU: the two-sided reduced solve (solver_kind 3; lba.cu: rev_gather_kernel, ldlt_win_kernel modes 1 / 2,
sep_merge_kernel) restated in numpy on the plan the product's host code makes (csrc/ldlt_plan.h through
lba_debug_two_sided_plan), against a dense solve -- for chain-like envelopes and for irregular ones (pose blocks
entering in uneven steps, as after a reverse Cuthill-McKee renumbering).  This holds the *scheme*: where the
separator lies, that the two sides never touch each other's columns, that side 1's window only holds its update when
the separator block starts from zero, the merge, and the back-substitution from the separator.  The kernels
themselves are held against the oracle on the GPU (tests/test_lba_gpu.py)."""
import ctypes as C

import numpy as np
import pytest

from orb_slam3_b200 import _lib


def _plan(reach):
    n = len(reach)
    r = np.ascontiguousarray(reach, np.int32)
    out = np.zeros(9, np.int32)
    f1, r1 = np.zeros(n, np.int32), np.zeros(n, np.int32)
    rc = _lib.lib().lba_debug_two_sided_plan(n, r.ctypes.data, out.ctypes.data, f1.ctypes.data, r1.ctypes.data)
    assert rc == 0
    keys = ("ok", "m", "e2", "p0", "p1", "w", "R0", "R1", "win_rows")
    return dict(zip(keys, [int(v) for v in out])), f1, r1


def _envelope(n_pose, rng, width_lo, width_hi, irregular):
    """first[i] per scalar row (multiples of 6, as lba_solve builds them from the covisibility pattern) and reach[c]."""
    bfirst = np.zeros(n_pose, int)
    for i in range(n_pose):
        wdt = rng.integers(width_lo, width_hi + 1)
        bfirst[i] = max(0, i - wdt)
        if irregular and rng.random() < 0.3:
            bfirst[i] = max(0, i - rng.integers(1, width_hi + 1))
    n = 6 * n_pose
    first = np.repeat(6 * bfirst, 6)
    reach = np.zeros(n, int)
    for i in range(n):
        reach[first[i]] = max(reach[first[i]], i)
    reach = np.maximum.accumulate(reach)
    return first, reach


def _spd_in_envelope(first, rng):
    n = len(first)
    A = np.zeros((n, n))
    for i in range(n):
        A[i, first[i]:i] = rng.normal(0, 1, i - first[i])
    A = A + A.T
    A += np.diag(np.abs(A).sum(1) + 1.0)   # diagonally dominant: positive definite, no pivoting needed
    return A


def _eliminate(S, b, cols):
    """Right-looking LDL^T over `cols` (in order) of a dense symmetric copy; returns the updated copies and L, D."""
    S, b = S.copy(), b.copy()
    n = len(b)
    L = np.zeros((n, n))
    D = np.zeros(n)
    z = np.zeros(n)
    for k in cols:
        D[k] = S[k, k]
        l = S[:, k] / D[k]
        l[k] = 0.0
        done = np.zeros(n, bool)
        done[list(cols[:list(cols).index(k) + 1])] = True
        l[done] = 0.0
        L[:, k] = l
        z[k] = b[k] / D[k]
        S -= np.outer(l, l) * D[k]
        b -= l * b[k]
    return S, b, L, D, z


@pytest.mark.parametrize("n_pose,lo,hi,irregular,seed", [(200, 10, 16, False, 0), (120, 6, 14, True, 1), (70, 8, 15, True, 2),
                                                         (260, 12, 16, True, 3), (90, 3, 6, False, 4)])
def test_two_sided_scheme_equals_a_dense_solve(n_pose, lo, hi, irregular, seed):
    rng = np.random.default_rng(seed)
    first, reach = _envelope(n_pose, rng, lo, hi, irregular)
    n = len(first)
    P, first1, reach1 = _plan(reach)
    assert P["ok"], P
    m, e2, w = P["m"], P["e2"], P["w"]
    # ---- the plan's invariants
    assert m % 8 == 0 and (n - e2) % 8 == 0 and m == 8 * P["p0"] and n - e2 == 8 * P["p1"] and w == e2 - m
    assert 8 <= w <= P["win_rows"] - 8
    assert (first[e2:] >= m).all(), "a row of side 1 reaches a column of side 0"
    assert P["R0"] < e2 and P["R0"] >= m - 1
    m1, e1 = n - e2, n - m
    assert P["R1"] <= e1 - 1
    assert abs(P["p0"] - P["p1"]) <= max(2, (w + 7) // 8), "the sides are balanced to within the separator's width"
    # side 1's tables are the envelope of the reversed matrix
    assert (first1 == n - 1 - reach[::-1]).all()
    A = _spd_in_envelope(first, rng)
    rev = A[::-1, ::-1]
    for a in range(n):
        assert not rev[a, :first1[a]].any(), "P S P leaves its envelope"
    b = rng.normal(0, 1, n)
    x_ref = np.linalg.solve(A, b)
    # ---- side 0: columns [0, m) top-down, in place
    S0, b0, L0, D0, z0 = _eliminate(A, b, list(range(m)))
    # its updates stay inside rows / columns < e2 (what mode 1 dumps: rows [m, R0])
    delta0 = S0 - A
    assert not delta0[e2:, :].any() and not delta0[P["R0"] + 1:, m:].any()
    # ---- side 1: the reversed matrix with the separator block and the separator's rhs entries zeroed
    M1 = rev.copy()
    M1[m1:e1, m1:e1] = 0.0
    b1 = b[::-1].copy()
    b1[m1:e1] = 0.0
    S1, bb1, L1, D1, z1 = _eliminate(M1, b1, list(range(m1)))
    upd = S1[m1:e1, m1:e1]                      # = -Delta_B in reversed numbering
    assert not S1[P["R1"] + 1:e1, m1:e1].any(), "side 1's update reaches beyond the rows its window holds"
    # ---- separator system: side 0's window + side 1's update (transposed back)
    Ssep = S0[m:e2, m:e2] + upd[::-1, ::-1]
    bsep = b0[m:e2] + bb1[m1:e1][::-1]
    xs = np.linalg.solve(Ssep, bsep)
    # ---- back-substitution of both sides from the separator (mode 2)
    x = np.zeros(n)
    x[m:e2] = xs
    acc = z0.copy()
    acc[m:e2] = xs
    for k in range(m - 1, -1, -1):
        x[k] = acc[k] - L0[k + 1:e2, k] @ x[k + 1:e2]
    xr = np.zeros(n)
    xr[m1:e1] = xs[::-1]
    for k in range(m1 - 1, -1, -1):
        xr[k] = z1[k] - L1[k + 1:e1, k] @ xr[k + 1:e1]
    x[e2:] = xr[:m1][::-1]
    assert np.abs(x - x_ref).max() <= 1e-9 * max(1.0, np.abs(x_ref).max())


def test_no_separator_means_no_plan():
    # every pose covisible with every other: the envelope is the whole triangle, there is nothing to split
    n = 6 * 20
    reach = np.full(n, n - 1)
    P, _, _ = _plan(reach)
    assert not P["ok"]
    # an envelope wider than a window
    first, reach = _envelope(120, np.random.default_rng(0), 25, 30, False)
    P, _, _ = _plan(reach)
    assert not P["ok"]

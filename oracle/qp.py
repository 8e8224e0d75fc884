"""Exact solver for GEM's dual QP (TEST INFRASTRUCTURE ONLY -- the oracle for a16).

Reference call site: src/methods/rehearsal/model/gem.py:70-79
    v = quadprog.solve_qp(P, q, G=I, h=margin*1)[0]
i.e. quadprog's  min 1/2 x^T G x - a^T x  s.t.  C^T x >= b  with C = I.

quadprog 0.1.6 (requirements.txt:51) is an un-vendored third-party dependency that is
not installed here.  The problem is strictly convex (P + 1e-3 I, gem.py:74) so the
minimiser is unique and any exact method is a valid oracle.  We enumerate the <= 2^k
active sets (k <= 9 tasks), solve each KKT system in fp64 and keep the candidate that is
primal- and dual-feasible.  Cross-checked against scipy (tests/test_qp.py).
"Parity unpinned": the reference holds no known-answer vectors for this boundary.
"""
import itertools

import numpy as np


def solve_lower_bounded_qp(G, a, lb):
    """min 1/2 x^T G x - a^T x  s.t. x >= lb   (G SPD).  Returns x (fp64)."""
    G = np.asarray(G, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    lb = np.asarray(lb, dtype=np.float64).reshape(-1)
    k = a.shape[0]
    assert k <= 16, "active-set enumeration is meant for GEM's k <= 9"
    scale = max(1.0, float(np.abs(G).max()), float(np.abs(a).max()))
    tol = 1e-9 * scale
    best, best_obj = None, np.inf
    fallback, fallback_viol = None, np.inf
    idx = np.arange(k)
    for nact in range(k + 1):
        for S in itertools.combinations(range(k), nact):
            S = np.array(S, dtype=np.int64)
            F = np.setdiff1d(idx, S)
            x = lb.copy()
            if F.size:
                rhs = a[F] - (G[np.ix_(F, S)] @ lb[S] if S.size else 0.0)
                x[F] = np.linalg.solve(G[np.ix_(F, F)], rhs)
            lam = G @ x - a          # multipliers of the active bounds (must be >= 0 on S)
            pv = float(np.maximum(lb[F] - x[F], 0).max()) if F.size else 0.0
            dv = float(np.maximum(-lam[S], 0).max()) if S.size else 0.0
            viol = max(pv, dv)
            if viol <= tol:
                obj = 0.5 * x @ G @ x - a @ x
                if obj < best_obj:
                    best, best_obj = x, obj
            if viol < fallback_viol:
                fallback, fallback_viol = x, viol
    return best if best is not None else fallback


def solve_qp_quadprog_signature(G, a, C=None, b=None, meq=0):
    """Drop-in for quadprog.solve_qp for the only form gem.py uses (C = I, meq = 0)."""
    k = np.asarray(a).shape[0]
    if C is None:
        raise ValueError("unconstrained form not needed by gem.py")
    C = np.asarray(C, dtype=np.float64)
    assert meq == 0 and C.shape == (k, k) and np.array_equal(C, np.eye(k)), \
        "oracle QP only covers gem.py's C = I inequality form"
    x = solve_lower_bounded_qp(G, a, b)
    obj = 0.5 * x @ np.asarray(G, dtype=np.float64) @ x - np.asarray(a, dtype=np.float64) @ x
    return (x, obj, None, None, None, None)


def project2cone2(gradient, memories, margin=0.5, eps=1e-3):
    """Restatement of gem.py:58-80 on numpy arrays.

    gradient: [P] fp32, memories: [k, P] fp32 (row i = memory gradient of past task i).
    Returns (x fp32 [P], v fp64 [k]).
    """
    M = np.asarray(memories, dtype=np.float64)
    g = np.asarray(gradient, dtype=np.float64).reshape(-1)
    t = M.shape[0]
    P = M @ M.T
    P = 0.5 * (P + P.T) + np.eye(t) * eps
    q = -(M @ g)
    v = solve_lower_bounded_qp(P, q, np.zeros(t) + margin)
    x = v @ M + g
    return x.astype(np.float32), v

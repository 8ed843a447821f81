"""Independent derivation of Hamilton's equations with sympy — pins the oracle as far as it can be pinned
without the Haskell binary.  TEST INFRASTRUCTURE ONLY.

Route (deliberately different from both the reference's formula and the oracle's jets):
  x = f(q) symbolic (the same number-polymorphic Python definitions the product traces, evaluated on sympy symbols),
  J = dx/dq by sympy.diff,  M(q) = J^T W J,
  H(q,p) = 1/2 p^T M^-1 p + U(q),   dq/dt = M^-1 p,
  dp_j/dt = -dH/dq_j = +1/2 v^T (dM/dq_j) v - dU/dq_j   with v = M^-1 p   (d(M^-1) = -M^-1 dM M^-1),
with dM/dq_j differentiated symbolically and everything evaluated in numpy (no symbolic inverse, so n = 12 works).
"""
import numpy as np
import sympy as sp


class SympySystem:
    def __init__(self, definition):
        inertia, f, u, n, cart = definition
        self.n, self.m = n, len(inertia)
        q = sp.symbols("q0:%d" % n, real=True)
        x = [sp.sympify(e) for e in f(list(q))]
        U = sp.sympify(u(x) if cart else u(list(q)))
        W = sp.diag(*[sp.Float(w) if not isinstance(w, int) else sp.Integer(w) for w in inertia])
        J = sp.Matrix(x).jacobian(sp.Matrix(q))
        M = J.T * W * J
        self._x = sp.lambdify([q], x, "numpy")
        self._J = sp.lambdify([q], J, "numpy")
        self._M = sp.lambdify([q], M, "numpy")
        self._dM = [sp.lambdify([q], M.diff(qj), "numpy") for qj in q]
        self._U = sp.lambdify([q], U, "numpy")
        self._gU = sp.lambdify([q], [sp.diff(U, qj) for qj in q], "numpy")

    def mass(self, q):
        return np.array(self._M(tuple(q)), dtype=float)

    def ham_eqs(self, q, p):
        q, p = tuple(np.asarray(q, float)), np.asarray(p, float)
        M = np.array(self._M(q), dtype=float)
        v = np.linalg.solve(M, p)
        gU = np.array(self._gU(q), dtype=float)
        dp = np.array([0.5 * v @ np.array(dM(q), dtype=float) @ v for dM in self._dM]) - gU
        return v, dp

    def hamiltonian(self, q, p):
        q, p = tuple(np.asarray(q, float)), np.asarray(p, float)
        return 0.5 * p @ np.linalg.solve(np.array(self._M(q), dtype=float), p) + float(self._U(q))

    def rhs(self, t, y):
        dq, dp = self.ham_eqs(y[: self.n], y[self.n:])
        return np.r_[dq, dp]

    def integrate(self, y0, t1, rtol=1e-12, atol=1e-13):
        from scipy.integrate import solve_ivp
        r = solve_ivp(self.rhs, (0.0, t1), np.asarray(y0, float), method="DOP853", rtol=rtol, atol=atol)
        return r.y[:, -1]

"""Second, independently written restatement of the integrator the reference runs (test infrastructure only).

`evolveHam` calls `odeSolveV RKf45 hi eps eps` (src/Numeric/Hamilton.hs:445-448), i.e. hmatrix-gsl's loop
`while t < t_i: gsl_odeiv2_evolve_apply(...)` over GSL's rkf45 stepper with the standard controller (a_y = a_dydt = 1).
This file restates GSL 2.x `ode-initval2/rkf45.c`, `cstd.c` and `evolve.c` in plain Python, from the published algorithm,
without looking at oracle/hamilton_oracle.c, so that a transcription slip in either (a tableau entry, the accept /
reject / FSAL logic, the step-size memory across output times) shows up as a disagreement
(tests/test_cpu_oracle.py::test_rkf45_two_restatements_agree).  The right-hand side is passed in as a callable."""
import math
import sys

import numpy as np

AH = (1 / 4, 3 / 8, 12 / 13, 1.0, 1 / 2)
B3 = (3 / 32, 9 / 32)
B4 = (1932 / 2197, -7200 / 2197, 7296 / 2197)
B5 = (8341 / 4104, -32832 / 4104, 29440 / 4104, -845 / 4104)
B6 = (-6080 / 20520, 41040 / 20520, -28352 / 20520, 9295 / 20520, -5643 / 20520)
C1, C3, C4, C5, C6 = 902880 / 7618050, 3953664 / 7618050, 3855735 / 7618050, -1371249 / 7618050, 277020 / 7618050
EC1, EC3, EC4, EC5, EC6 = 1 / 360, -128 / 4275, -2197 / 75240, 1 / 50, 2 / 55
EPS = 1.49012e-08                    # src/Numeric/Hamilton.hs:448 (both eps_abs and eps_rel)
HADJ_DEC, HADJ_NIL, HADJ_INC = -1, 0, 1


class Stats:
    def __init__(self):
        self.steps = self.rejects = self.rhs_evals = 0


def rkf45_apply(f, h, y, k1, st):
    """One step: returns (y_new, yerr, dydt_out); the 5th-order solution is the one advanced."""
    k2 = f(y + AH[0] * h * k1)
    k3 = f(y + h * (B3[0] * k1 + B3[1] * k2))
    k4 = f(y + h * (B4[0] * k1 + B4[1] * k2 + B4[2] * k3))
    k5 = f(y + h * (B5[0] * k1 + B5[1] * k2 + B5[2] * k3 + B5[3] * k4))
    k6 = f(y + h * (B6[0] * k1 + B6[1] * k2 + B6[2] * k3 + B6[3] * k4 + B6[4] * k5))
    y_new = y + h * (C1 * k1 + C3 * k3 + C4 * k4 + C5 * k5 + C6 * k6)
    yerr = h * (EC1 * k1 + EC3 * k3 + EC4 * k4 + EC5 * k5 + EC6 * k6)
    dydt_out = f(y_new)
    st.rhs_evals += 6
    return y_new, yerr, dydt_out


def std_control_hadjust(y, yerr, yp, h_old, order=5):
    S = 0.9
    rmax = sys.float_info.min
    for i in range(len(y)):
        d0 = EPS * (1.0 * abs(y[i]) + 1.0 * abs(h_old * yp[i])) + EPS
        r = abs(yerr[i]) / abs(d0)
        rmax = max(r, rmax)
    if rmax > 1.1:
        r = S / math.pow(rmax, 1.0 / order)
        if r < 0.2:
            r = 0.2
        return r * h_old, HADJ_DEC
    if rmax < 0.5:
        r = S / math.pow(rmax, 1.0 / (order + 1.0))
        if r > 5.0:
            r = 5.0
        if r < 1.0:
            r = 1.0
        return r * h_old, HADJ_INC
    return h_old, HADJ_NIL


class Evolve:
    """gsl_odeiv2_evolve: remembers dydt_out of the last accepted step (count > 0) across calls."""

    def __init__(self):
        self.count = 0
        self.dydt_out = None

    def apply(self, f, t, t1, h, y, st):
        """One evolve_apply: returns (t, h, y) after ONE accepted step (retrying rejected ones)."""
        t0, h0, dt = t, h, t1 - t
        y0 = y.copy()
        if self.count == 0:
            dydt_in = f(y0)
            st.rhs_evals += 1
        else:
            dydt_in = self.dydt_out.copy()
        while True:
            if (dt >= 0.0 and h0 > dt) or (dt < 0.0 and h0 < dt):
                h0, final_step = dt, True
            else:
                final_step = False
            y, yerr, dydt_out = rkf45_apply(f, h0, y0, dydt_in, st)
            self.count += 1
            t = t1 if final_step else t0 + h0
            h_old = h0
            h0, status = std_control_hadjust(y, yerr, dydt_out, h_old)
            if status == HADJ_DEC:
                t_curr, t_next = t, t + h0
                if abs(h0) < abs(h_old) and t_next != t_curr:
                    st.rejects += 1          # undo and retry with the smaller step
                    continue
                raise ArithmeticError("step size cannot be decreased")
            break
        self.dydt_out = dydt_out
        st.steps += 1
        if not final_step:
            h = h0                            # no suggestion from the (possibly tiny) last step of an interval
        return t, h, y


def ode_solve(f, y0, ts):
    """hmatrix-gsl odeSolveV: rows at every ts[i]; h starts at (ts[1]-ts[0])/100 (src/Numeric/Hamilton.hs:447)."""
    y = np.array(y0, dtype=float)
    st, ev = Stats(), Evolve()
    t, h = float(ts[0]), (float(ts[1]) - float(ts[0])) / 100
    rows = [y.copy()]
    for ti in ts[1:]:
        while t < ti:
            t, h, y = ev.apply(f, t, float(ti), h, y, st)
        rows.append(y.copy())
    return np.array(rows), st

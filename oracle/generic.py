"""TEST INFRASTRUCTURE -- CPU restatement of the reference's generic Hamiltonian / partial functions.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this package; the product (levelsetpy_b200/) never does.

Follows Hamiltonians/generic_ham.py:5-57 and Hamiltonians/generic_partial.py:6-58 line by line (numpy), for the fields the
device path supports (no uIn / dIn / deriv / side / TIdim / dynSys.partialFunc).  ``sd.dynSys`` is any object with the
reference's dynSys API: get_opt_u(t, deriv, uMode, y), get_opt_v(t, deriv, dMode, y), dynamics(t, x, u, d), nx.

Pinned against the LITERAL reference by tests/golden/make_golden_generic.py (bit for bit) -> tests/golden/generic_dyn.npz.
"""
import numpy as np


def generic_ham(t, data, deriv, sd):
    dyn = sd.dynSys                                                        # generic_ham.py:8
    if "uMode" not in sd.__dict__:
        sd.uMode = "min"                                                   # :10-11
    if "dMode" not in sd.__dict__:
        sd.dMode = "max"                                                   # :13-14
    if "tMode" not in sd.__dict__:
        sd.tMode = "backward"                                              # :16-17
    u = dyn.get_opt_u(t, deriv, uMode=sd.uMode, y=sd.grid.xs)              # :27
    d = dyn.get_opt_v(t, deriv, dMode=sd.dMode, y=sd.grid.xs)              # :32
    ham = 0
    dx = dyn.dynamics(t, sd.grid.xs, u, d)                                 # :45
    for i in range(dyn.nx):
        ham += deriv[i] * dx[i]                                            # :46-47
    if sd.tMode == "backward":
        ham = -ham                                                         # :54-55
    return ham


def generic_partial(t, data, deriv_min, deriv_max, sd, dim):
    g = sd.grid
    dyn = sd.dynSys
    if "uMode" not in sd.__dict__:
        sd.uMode = "min"                                                   # generic_partial.py:16-17
    if "dMode" not in sd.__dict__:
        sd.dMode = "min"                                                   # :19-20 (sic: 'min' here, 'max' in genericHam)
    uU = dyn.get_opt_u(t, deriv_max, sd.uMode, g.xs)                       # :28
    uL = dyn.get_opt_u(t, deriv_min, sd.uMode, g.xs)                       # :31
    dU = dyn.get_opt_v(t, deriv_max, sd.dMode, g.xs)                       # :39
    dL = dyn.get_opt_v(t, deriv_min, sd.dMode, g.xs)                       # :40
    dxUU = dyn.dynamics(t, g.xs, uU, dU)                                   # :43-46
    dxUL = dyn.dynamics(t, g.xs, uU, dL)
    dxLL = dyn.dynamics(t, g.xs, uL, dL)
    dxLU = dyn.dynamics(t, g.xs, uL, dU)
    alpha = np.maximum(np.abs(dxUU[dim]), np.abs(dxUL[dim]))               # :49
    alpha = np.maximum(alpha, np.abs(dxLL[dim]))                           # :50
    alpha = np.maximum(alpha, np.abs(dxLU[dim]))                           # :51
    return alpha


class DubinsCar:
    """Test dynSys with the reference's dynSys API (what a user of genericHam writes): the Dubins car with disturbances.
    Plain numpy, independent of the product's class of the same name."""

    def __init__(self, speed, wMax, dMax):
        self.nx = 3
        self.speed, self.wMax, self.dMax = float(speed), float(wMax), [float(v) for v in dMax]

    def get_opt_u(self, t, deriv, uMode="min", y=None):
        if uMode == "max":
            return (deriv[2] >= 0) * self.wMax + (deriv[2] < 0) * (-self.wMax)
        return (deriv[2] >= 0) * (-self.wMax) + (deriv[2] < 0) * self.wMax

    def get_opt_v(self, t, deriv, dMode="max", y=None):
        out = []
        for i in range(3):
            if dMode == "max":
                out.append((deriv[i] >= 0) * self.dMax[i] + (deriv[i] < 0) * (-self.dMax[i]))
            else:
                out.append((deriv[i] >= 0) * (-self.dMax[i]) + (deriv[i] < 0) * self.dMax[i])
        return out

    def dynamics(self, t, x, u, d):
        return [self.speed * np.cos(x[2]) + d[0], self.speed * np.sin(x[2]) + d[1], u + d[2]]

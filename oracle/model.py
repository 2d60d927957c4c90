"""Vessel model + integrator restated (FP64).  TEST INFRASTRUCTURE.

Follows (file:line under /root/reference/gym_auv):
  utils/constants.py:4-6,12-13 (m, x_g, I_z, added masses), :33-37 (M, M_inv),
  :39-43 (D), :63-72 (N(nu));  utils/geomutils.py:4-5 (princip), :37-43 (Rz);
  objects/vessel/odesolver.py:2-47 (Fehlberg 4(5) single step, returns (w4, q5));
  objects/vessel/vessel.py:561-570 (_state_dot), :572-578 (input clipping),
  utils/sector_partitioning.py:4-9 (sector map).
PINNED: checked against the reference's own files by
``tests/golden/make_reference_goldens.py`` / ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

import math

import numpy as np

_m, _xg, _Iz = 23.8, 0.046, 1.760
_Xud, _Yvd, _Yrd, _Nrd, _Nvd = -2.0, -10.0, 0.0, -1.0, 0.0
_Xu, _Yv, _Yr, _Nv, _Nr = -2.0, -7.0, -0.1, -0.1, -0.5

MASS = np.array(
    [
        [_m - _Xud, 0.0, 0.0],
        [0.0, _m - _Yvd, _m * _xg - _Yrd],
        [0.0, _m * _xg - _Nvd, _Iz - _Nrd],
    ]
)
MASS_INV = np.linalg.inv(MASS)
DAMP = np.array([[2.0, 0.0, 0.0], [0.0, 7.0, -2.5425], [0.0, -2.5425, 1.422]])


def nonlinear_damping(nu):
    u = nu[0]
    return np.array(
        [
            [-_Xu, 0.0, 0.0],
            [0.0, -_Yv, _m * u - _Yr],
            [0.0, -_Nv, _m * _xg * u - _Nr],
        ]
    )


def princip(angle):
    """Python floored modulo => result in [-pi, pi)."""
    return ((angle + math.pi) % (2.0 * math.pi)) - math.pi


def state_dot(state, tau_u, tau_r):
    psi = princip(state[2])
    nu = state[3:6]
    c, s = math.cos(psi), math.sin(psi)
    eta_dot = np.array([c * nu[0] - s * nu[1], s * nu[0] + c * nu[1], nu[2]])
    tau = np.array([tau_u, 0.0, tau_r])
    nu_dot = MASS_INV.dot(tau - DAMP.dot(nu) - nonlinear_damping(nu).dot(nu))
    return np.concatenate([eta_dot, nu_dot])


def rkf45_step(f, y, h):
    """One Fehlberg step; returns (4th-order w, 5th-order q).  The caller keeps q."""
    k1 = f(y)
    k2 = f(y + h * k1 / 4.0)
    k3 = f(y + 3.0 * h * k1 / 32.0 + 9.0 * h * k2 / 32.0)
    k4 = f(y + 1932.0 * h * k1 / 2197.0 - 7200.0 * h * k2 / 2197.0 + 7296.0 * h * k3 / 2197.0)
    k5 = f(y + 439.0 * h * k1 / 216.0 - 8.0 * h * k2 + 3680.0 * h * k3 / 513.0 - 845.0 * h * k4 / 4104.0)
    k6 = f(
        y
        - 8.0 * h * k1 / 27.0
        + 2 * h * k2
        - 3544.0 * h * k3 / 2565
        + 1859.0 * h * k4 / 4104.0
        - 11.0 * h * k5 / 40.0
    )
    w = y + h * (25.0 * k1 / 216.0 + 1408.0 * k3 / 2565.0 + 2197.0 * k4 / 4104.0 - k5 / 5.0)
    q = y + h * (
        16.0 * k1 / 135.0 + 6656.0 * k3 / 12825.0 + 28561.0 * k4 / 56430.0 - 9.0 * k5 / 50.0 + 2.0 * k6 / 55.0
    )
    return w, q


def vessel_step(state, action, dt, thrust_max=2.0, moment_max=0.15):
    """``Vessel.step`` without the history bookkeeping (vessel.py:226-247)."""
    tau_u = float(np.clip(action[0], 0.0, 1.0)) * thrust_max
    tau_r = float(np.clip(action[1], -1.0, 1.0)) * moment_max
    _, q = rkf45_step(lambda s: state_dot(s, tau_u, tau_r), np.asarray(state, dtype=np.float64), dt)
    q = np.array(q, dtype=np.float64)
    q[2] = princip(q[2])
    return q


def sector_of_ray(isensor, n_sensors=180, n_sectors=9, c=0.1):
    """utils/sector_partitioning.py:4-9."""
    a, b = n_sensors, n_sectors

    def sigma(x):
        return b / (1 + np.exp((-x + a / 2) / (c * a)))

    return int(np.floor(sigma(isensor) - sigma(0)))


def sector_table(n_sensors=180, n_sectors=9, c=0.1):
    return np.array([sector_of_ray(i, n_sensors, n_sectors, c) for i in range(n_sensors)], dtype=np.int32)

"""Oracle: ODE steppers used to drive the golden trajectories.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  rhs is called as f(y, t, intStep=k).
"""
import numpy as np


def euler_step(f, y, t, dt):
    """jVMC/util/stepper.py:14-40."""
    return y + dt * f(y, t, intStep=0), dt


def heun_step(f, y, t, dt):
    """jVMC/util/stepper.py:62-93."""
    k0 = f(y, t, intStep=0)
    k1 = f(y + dt * k0, t + dt, intStep=1)
    return y + 0.5 * dt * (k0 + k1), dt


class AdaptiveHeun:
    """jVMC/util/stepper.py:98-182: one full Heun step vs two half steps (5 rhs calls per
    attempt), step size rescaled by clip(0.9*fe^(1/3), 0.2, 2) until fe = tol/err >= 1."""

    def __init__(self, timeStep=1e-3, tol=1e-8, maxStep=1.0):
        self.dt, self.tol, self.maxStep = timeStep, tol, maxStep

    def step(self, f, y0, t, norm=np.linalg.norm):
        dt, fe = self.dt, 0.5
        while fe < 1.0:
            k0 = f(y0, t, intStep=0)
            k1 = f(y0 + dt * k0, t + dt, intStep=1)
            full = 0.5 * dt * (k0 + k1)
            k10 = f(y0 + 0.5 * dt * k0, t + 0.5 * dt, intStep=2)
            half = 0.25 * dt * (k0 + k10)
            ymid = y0 + half
            k01 = f(ymid, t + 0.5 * dt, intStep=3)
            k11 = f(ymid + 0.5 * dt * k01, t + dt, intStep=4)
            half = half + 0.25 * dt * (k01 + k11)
            fe = self.tol / norm(half - full)
            used = dt
            dt = min(dt * min(max(0.9 * fe ** 0.33333, 0.2), 2.0), self.maxStep)
        self.dt = dt
        return y0 + half, used

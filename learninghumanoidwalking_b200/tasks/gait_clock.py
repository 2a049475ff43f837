"""Gait phase clocks as integer-phase lookup tables.

Restates tasks/rewards.py:196-300 (create_phase_reward) of the reference: 8 knots per cycle (swing /
double-stance plateaus at +-1, shrunk by `strict_relaxer`), repeated over three cycles, PCHIP-interpolated.
The reference rebuilds the four splines with scipy on every task.reset() (tasks/walking_task.py:198-200)
although they depend only on YAML constants and are only ever evaluated at integer phases; here they are
tabulated once and uploaded as a 4 x period constant table (r_frc, r_vel, l_frc, l_vel).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import PchipInterpolator


def phase_clock_table(swing_duration: float, stance_duration: float, strict_relaxer: float = 0.1,
                      stance_mode: str = "grounded", freq: float = 40.0, total_duration: float | None = None):
    """Returns (period, table[4, period]) with rows r_frc, r_vel, l_frc, l_vel."""
    sw, st = swing_duration * freq, stance_duration * freq
    # segment boundaries of one cycle: right swing | double stance | left swing | double stance
    bounds = np.array([0.0, sw, sw + st, 2 * sw + st, 2 * (sw + st)])
    xs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        off = (b - a) * strict_relaxer
        xs += [a + off, b - off]
    xs = np.array(xs)
    last_off = (bounds[4] - bounds[3]) * strict_relaxer
    stance_val = {"grounded": 1.0, "aerial": -1.0, "zero": 0.0}[stance_mode]
    # right foot force clock per segment: swinging (-1), stance, supporting (+1), stance
    r_frc = np.array([-1, -1, stance_val, stance_val, 1, 1, stance_val, stance_val], dtype=float)
    l_frc = np.array([1, 1, stance_val, stance_val, -1, -1, stance_val, stance_val], dtype=float)
    shift = xs[-1] + last_off
    x3 = np.concatenate([xs - shift, xs, xs + shift])

    def spline(y):
        return PchipInterpolator(x3, np.tile(y, 3))

    period = int(np.floor(2 * (total_duration if total_duration is not None else (swing_duration + stance_duration)) * freq))
    ph = np.arange(period)
    table = np.stack([spline(r_frc)(ph), spline(-r_frc)(ph), spline(l_frc)(ph), spline(-l_frc)(ph)])
    return period, table

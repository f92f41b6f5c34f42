"""Oracle: Gamma-process loss-trajectory simulation (reference autotune/core/simulation_evaluator.py).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Scalar Python-float loops in the reference's operation order, so a
trajectory computed here from the same Gamma draws is bit-identical to the reference's `non_smooth_fs`.
"""
import numpy as np
from scipy.signal import savgol_filter

K_MODE = 2  # simulation_evaluator.py:94 - mode of every Gamma distribution, the "neutral" aggressiveness


def gamma_shape_scale(t, R):
    """(shape, scale) of the draw made at step t of an R-point trajectory.

    simulation_evaluator.py:36-42 with the call site's arguments (time=t, n=R+1, k=2) (simulation_evaluator.py:98).
    """
    m = (R + 1) - t
    root = np.sqrt(K_MODE ** 2 + 4 * m)
    beta = (K_MODE + root) / (2 * m)
    alpha = K_MODE * beta + 1
    return alpha, 1 / beta


def step(f, g, t, R, f_n, ml_agg, nec_agg, up_spik):
    """One epoch t -> t+1 of the recurrence (simulation_evaluator.py:99-115 with n = R)."""
    n = R
    if g == K_MODE:                                    # :100-101
        return f
    if g > K_MODE:                                     # :102-107
        ml = f + (g - K_MODE) * ml_agg * (f_n - f) / 100
        return ml + (f_n - ml) * ((t / (n - 1)) ** nec_agg)
    up = f + up_spik * 1 / (1 + g)                     # :108-115
    nxt = up + (f_n - up) * ((t / (n - 1)) ** (1.1 * nec_agg))
    if n - t == 1:
        nxt = f_n
    return nxt


def simulate(f_1, f_n, R, ml_agg, nec_agg, up_spik, draw):
    """Whole raw trajectory [f_1 .. f_R]; `draw(t)` supplies the Gamma value of step t = 1..R-1
    (simulation_evaluator.py:84-116 called as simulate(R, R, f_n) on non_smooth_fs == [f_1])."""
    fs = [f_1]
    for t in range(1, R):
        fs.append(step(fs[-1], draw(t), t, R, f_n, ml_agg, nec_agg, up_spik))
    return fs


def savgol_window(R):
    """simulation_evaluator.py:80-81."""
    w = int(0.17 * R + 6)
    return w + 1 if w % 2 == 0 else w


def readout(raw, is_smooth, R):
    """The `.fs` property (simulation_evaluator.py:76-82): raw list, or cubic Savitzky-Golay over the whole trajectory.
    Raises ValueError like scipy when the window exceeds R (e.g. R=6)."""
    if not is_smooth:
        return raw
    return list(savgol_filter(raw, savgol_window(R), 3))


def savgol_operator(R):
    """The R x R matrix W with savgol_filter(x, window, 3) == W @ x (the filter is linear; SURVEY 7 hard part 4).
    Used to give device kernels the few rows they need; agreement with scipy's own path is ~1e-13, not bit-exact."""
    w = savgol_window(R)
    if w > R:
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    return savgol_filter(np.eye(R), w, 3, axis=0)

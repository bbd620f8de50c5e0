"""Error metric shared by the parity tests.

rel_i = |x_i - ref_i| / (|ref_i| + 1e-3 * nominal_i), max over the 12 state components per aircraft.
`nominal_i` is a fixed in-envelope magnitude per component (feet, radians, ft/s, rad/s), so that components which
are ~0 right after a reset (beta, P, R ...) are measured against a physically meaningful floor instead of against
their own rounding noise.
"""
import numpy as np

#                     npos  epos   alt   phi theta  psi    vt  alpha beta   P    Q    R
NOMINAL = np.array([1e3, 1e3, 2e4, 0.5, 0.2, 1.0, 1e3, 0.2, 0.1, 0.5, 0.2, 0.2], dtype=np.float64)
FLOOR = 1e-3 * NOMINAL


def state_rel_err(s, s_ref):
    """[m,12] -> [m] max-component relative error."""
    s, s_ref = np.asarray(s, dtype=np.float64), np.asarray(s_ref, dtype=np.float64)
    return (np.abs(s - s_ref) / (np.abs(s_ref) + FLOOR)).max(axis=1)

"""Deterministic input tapes shared by the golden generator, the oracle tests and the GPU parity tests.

TEST INFRASTRUCTURE (see oracle/f16_oracle.py header).  A counter-based splitmix64 hash in
numpy uint64 arithmetic: no dependence on any library's RNG stream, so a tape regenerated on
another machine / numpy / torch version is bit-identical to the one the fixtures were made with.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def uniform01(seed, stream, shape):
    """float32 uniforms in [0, 1) with 24 random bits; (seed, stream) select an independent tape."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed], dtype=np.uint64) * np.uint64(0x632BE59BD9B4E019)
                           + np.array([stream], dtype=np.uint64))
        ctr = np.arange(n, dtype=np.uint64) + base
    z = _splitmix64(ctr)
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(shape)


def action_tape(seed, step, n, scale=1.0, num_actions=4):
    """Actions for one step: scale * U(-1, 1)^num_actions, float32 [n, num_actions]."""
    u = uniform01(seed, 2 * step, (n, num_actions))
    return (np.float32(scale) * (np.float32(2.0) * u - np.float32(1.0))).astype(np.float32)


def reset_draw_tape(seed, step, n, width=5):
    """Reset draws consumed by lanes that reset at the top of `step` (step 0 = the initial reset())."""
    return uniform01(seed, 2 * step + 1, (n, width))


def random_envelope_states(seed, n):
    """In-envelope (s[n,12], u[n,5]) samples for single-step / nlplant parity (SURVEY 8c protocol)."""
    r = uniform01(seed, 1_000_003, (n, 17))
    lo = np.array([-5e4, -5e4, 3000, -np.pi, -1.2, -np.pi, 300, -0.30, -0.45, -1.5, -1.0, -1.0,
                   -2000, -25, -21.5, -30, 0], dtype=np.float32)
    hi = np.array([5e4, 5e4, 40000, np.pi, 1.2, np.pi, 1500, 0.75, 0.45, 1.5, 1.0, 1.0,
                   19000, 25, 21.5, 30, 0], dtype=np.float32)
    x = (lo + (hi - lo) * r).astype(np.float32)
    return x[:, :12].copy(), x[:, 12:].copy()

"""Import the UNMODIFIED reference (xuecy22/NeuralPlane) on CPU, for golden generation only.

Test infrastructure.  Works only where /root/reference exists (the build
container); nothing in the product, the `-m gpu` tests, smoke() or bench.py
imports this file.  The reference needs two absent third-party modules, stubbed
under tests/golden/_shims (gym, torchdiffeq -- see SURVEY.md App. F).

Reset randomness is made injectable by patching `torch.rand_like` / `torch.rand`
while `env.reset()` runs, so the reference and the CUDA path consume the same
"reset-draw tape": tape[k, i, j] is the j-th uniform draw aircraft i would use
if it resets at the top of step k (j=0 altitude, j=1 vt  -- F16_model.py:41-42;
j=2.. task draws -- control_task.py:59-61, tracking_task.py:57-60).
"""
import contextlib
import io
import os
import sys

import torch

REF_ROOT = os.environ.get("NPLANE_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "envs"))


def _paths():
    for p in (os.path.join(REF_ROOT, "envs"), REF_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)


def import_reference():
    _paths()
    import envs.control_env as control_env  # noqa

    return control_env


class RefEnv:
    """Reference ControlEnv on CPU with silenced prints and injectable reset draws."""

    def __init__(self, num_envs, config="heading", model="F16", seed=0, noise_scale=0.0):
        ce = import_reference()
        with contextlib.redirect_stdout(io.StringIO()):
            self.env = ce.ControlEnv(num_envs=num_envs, config=config, model=model,
                                     random_seed=seed, device="cpu")
        self.env.task.noise_scale = noise_scale
        self.n = self.env.n
        self._draws = None
        self._col = 0
        self._mask = None

    # -- draw injection ---------------------------------------------------
    @contextlib.contextmanager
    def _inject(self, draws):
        """draws: [n, D] f32 uniforms in [0,1); lanes that reset consume columns in call order."""
        if draws is None:
            yield
            return
        env = self.env
        mask = (env.is_done.bool() | env.bad_done.bool()) | env.exceed_time_limit.bool()
        state = {"col": 0}
        orig_rand_like, orig_rand = torch.rand_like, torch.rand

        def rand_like(t, *a, **k):
            c = state["col"]
            state["col"] += 1
            out = draws[mask, c].to(t.dtype)
            assert out.shape == t.shape
            return out

        def rand(*size, **k):
            c = state["col"]
            state["col"] += 1
            out = draws[mask, c]
            return out

        torch.rand_like, torch.rand = rand_like, rand
        try:
            yield
        finally:
            torch.rand_like, torch.rand = orig_rand_like, orig_rand

    def reset(self, draws=None):
        with contextlib.redirect_stdout(io.StringIO()), self._inject(draws):
            return self.env.reset()

    def step(self, action, draws=None, perturb_after_reset=None):
        """One reference step; `draws` feeds the reset at the top of step().  `perturb_after_reset` (a factor such
        as 1 + 2^-23) scales the state of aircraft that have just been re-initialised: used to measure the
        reference's own sensitivity to a 1-ulp perturbation."""
        env = self.env
        with contextlib.redirect_stdout(io.StringIO()):
            # BaseEnv.step = reset(); update; count; obs; done; reward (env_base.py:99-109)
            mask = (env.is_done.bool() | env.bad_done.bool()) | env.exceed_time_limit.bool()
            with self._inject(draws):
                env.reset()
            if perturb_after_reset is not None:
                env.model.s[mask] = env.model.s[mask] * perturb_after_reset
            env.model.update(action)
            env.step_count += 1
            obs = env.obs()
            done, bad_done, exceed, _ = env.done({})
            reward = env.reward()
        return obs, reward, done, bad_done, exceed

    # -- state access -------------------------------------------------------
    def snapshot(self):
        e = self.env
        t = e.task
        names = [k for k in ("target_altitude", "target_heading", "target_vt", "target_pitch",
                             "target_npos", "target_epos") if hasattr(t, k)]
        return {
            "s": e.model.s.clone(), "u": e.model.u.clone(),
            "step_count": e.step_count.clone(),
            "is_done": e.is_done.clone().bool(), "bad_done": e.bad_done.clone().bool(),
            "exceed_time_limit": e.exceed_time_limit.clone().bool(),
            **{k: getattr(t, k).clone() for k in names},
        }

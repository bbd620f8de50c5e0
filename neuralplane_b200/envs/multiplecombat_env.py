"""MultipleCombatEnv (reference: envs/multiplecombat_env.py:25-274, envs/configs/multiple_selfplay.yaml): 2-v-2 combat,
four aircraft per env, one kernel launch per env step (np_env_combat_step with combat_pairs_per_env = 2).

The reference file is the 1-v-1 code with `num_agents: 4`: its obs / reward / blood / Crash / Shutdown pair only agents 0
and 1 of each env (`ego = e * num_agents`, `enm = ego + 1`: multiplecombat_env.py:100-101,165-166,260-261, crash.py:32-33,
shutdown.py:33-34), so its observation stack has 2 E relative rows against 4 E own-state rows and cannot be built; it
cannot be constructed either (same stale BaseEnv call as SingleCombatEnv).  The pairing is therefore RESTATED, with the
reference's own per-pair formulas and everything else as the file has it:

  * an env is two adjacent duels, agents [ego0, enm0, ego1, enm1]; duel d = (agent 2d, agent 2d + 1) gets the reference's
    1-v-1 geometry, observation mirroring, reward, blood model, Crash and Shutdown;
  * the env-level reset spans all four agents (reset_done_envs: `torch.any` over num_agents, :207-238);
  * ONE FDM step per env step (the reference calls super().step once, :258, against 5 in SingleCombatEnv);
  * reward = orientation_reward x range_reward without SingleCombatEnv's 0.01 factor (:163-181);
  * initial positions in +-10 000 ft (multiple_selfplay.yaml:35-38).

Parity: every formula is pinned through the 1-v-1 fixtures (tests/golden/combat*_traj.npz); the orchestration above is
checked against oracle/combat_oracle.py:MultiCombatOracle -- parity UNPINNED for the orchestration (no runnable reference).
"""
from .singlecombat_env import SingleCombatEnv


class MultipleCombatEnv(SingleCombatEnv):
    n_substeps = 1                      # multiplecombat_env.py:258
    combat_pairs_per_env = 2
    combat_reward_scale = 1.0           # multiplecombat_env.py:176-177

    def __init__(self, num_envs=1, config='multiple_selfplay', random_seed=None, device="cuda:0", **kw):
        if kw.get("layout", "pair") != "pair":
            raise NotImplementedError("MultipleCombatEnv keeps the four aircraft of an env on one rank (layout='pair')")
        super().__init__(num_envs, config, random_seed, device, **kw)
        if self.num_agents != 4:
            raise NotImplementedError("Number of agents must be equal to ego plus enm!")

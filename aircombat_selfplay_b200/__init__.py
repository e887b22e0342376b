"""aircombat_selfplay_b200 -- B200-native batched air-combat simulator: a drop-in for the env step of
junghoseong/aircombat-selfplay (SingleControlEnv / SingleCombatEnv / MultipleCombatEnv behind the VecEnv contract).

    from aircombat_selfplay_b200.env_wrappers import ShareSubprocVecEnv, SubprocVecEnv      # the reference's names
    from aircombat_selfplay_b200.envs import SingleControlEnv, SingleCombatEnv, MultipleCombatEnv, BatchedEnv

All device work goes through the C ABI of include/acs.h (``libacs.so``, built by ``__graft_entry__.build``); there is
no CPU fallback.
"""
__version__ = "0.1.0"

"""RLGPUEnv: rl_games IVecEnv adapter [ref: OIGE/utils/rlgames/rlgames_utils.py:102-126 ; RLG/common/ivecenv.py:1-36]."""
from __future__ import annotations


class RLGPUEnv:
    def __init__(self, env):
        self.env = env

    def step(self, action):
        return self.env.step(action)

    def reset(self):
        return self.env.reset()

    def get_number_of_agents(self):
        return 1

    def get_env_info(self):
        info = {"action_space": self.env.action_space, "observation_space": self.env.observation_space}
        if getattr(self.env, "num_states", 0) > 0:
            info["state_space"] = self.env.state_space
        return info

    def set_train_info(self, env_frames, *args, **kwargs):
        pass

    def get_env_state(self):
        return None

    def set_env_state(self, env_state):
        pass

"""PpoPlayerContinuous: rl_games' `test=True` path (policy evaluation) over the fused env and the CUDA policy kernels.

Mirrors the reference surface [ref: RLG/algos_torch/players.py:116-219 (PpoPlayerContinuous), RLG/common/player.py:147-207 (BasePlayer
constructor: the `player` section of the train YAML), :226-240 (env_step), :283-285 (env_reset), :319-422 (run)]:
`restore(fn)` reads the reference's `.pth` schema (`checkpoint["model"]`; written by `A2CAgent.save` or by rl_games itself),
`get_action(obs, is_deterministic)` returns mu or a sample, clamped to [-1, 1] and rescaled to the action-space bounds, and `run()`
plays `games_num` episodes and prints `av reward` / `av steps` with the reference's accounting.  Inference is the same sm_100a kernel the
training rollouts use (`PolicyMLP.act`): there is no torch module and no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from .policy import PolicyMLP

_NOT_WEIGHTS = ("epoch", "frame", "last_mean_rewards")


def rescale_actions(low: torch.Tensor, high: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """[ref: RLG/algos_torch/players.py:11-15]"""
    d = (high - low) / 2.0
    m = (high + low) / 2.0
    return action * d + m


class PpoPlayerContinuous:
    def __init__(self, vec_env, player_config: Optional[dict] = None, device="cuda:0", clip_actions: bool = True,
                 tensor_cores: bool = True, seed: int = 0):
        """`vec_env`: the rl_games IVecEnv (RLGPUEnv over VecEnvRLGames); `player_config`: the YAML's `params.config.player` section
        (`games_num`, `deterministic`, `n_game_life`, `print_stats`; defaults as in the reference)."""
        self.env = vec_env
        self.player_config = dict(player_config or {})
        self.env_info = vec_env.get_env_info()
        self.num_agents = int(self.env_info.get("agents", 1))
        self.value_size = int(self.env_info.get("value_size", 1))
        self.action_space = self.env_info["action_space"]
        self.observation_space = self.env_info["observation_space"]
        self.obs_shape = {k: v.shape for k, v in self.observation_space.spaces.items()} \
            if hasattr(self.observation_space, "spaces") else self.observation_space.shape
        self.device = torch.device(device)
        self.clip_actions = bool(clip_actions)
        self.games_num = int(self.player_config.get("games_num", 2000))
        self.is_deterministic = bool(self.player_config.get("deterministic", True))
        self.n_game_life = int(self.player_config.get("n_game_life", 1))
        self.print_stats = bool(self.player_config.get("print_stats", True))
        self.max_steps = 108000 // 4
        self.is_rnn, self.states, self.is_tensor_obses = False, None, True
        self.actions_num = int(self.action_space.shape[0])
        self.actions_low = torch.as_tensor(self.action_space.low, dtype=torch.float32).to(self.device)
        self.actions_high = torch.as_tensor(self.action_space.high, dtype=torch.float32).to(self.device)
        obs_dim = int(self.observation_space["state"].shape[0])
        self.model = PolicyMLP(obs_dim, self.device, seed=seed, tensor_cores=tensor_cores)

    # ---- weights  [ref: players.py:215-219, player.py:290-299] ---------------------------------------
    def restore(self, fn: str) -> None:
        checkpoint = torch.load(fn, map_location="cpu", weights_only=False)
        self.model.load_state_dict(checkpoint["model"])

    def get_weights(self) -> dict:
        return {"model": self.model.state_dict()}

    def set_weights(self, weights: dict) -> None:
        self.model.load_state_dict({k: torch.as_tensor(v) for k, v in weights["model"].items() if k not in _NOT_WEIGHTS})

    # ---- one step  [ref: players.py:172-213] ------------------------------------------------------------
    def get_action(self, obs, is_deterministic: bool = False) -> torch.Tensor:
        if isinstance(obs, dict):
            obs = obs["obs"] if "obs" in obs else obs
            obs = obs["state"] if isinstance(obs, dict) else obs
        res = self.model.act(obs.to(self.device))
        current_action = res["mus"] if is_deterministic else res["actions"]
        if self.clip_actions:
            return rescale_actions(self.actions_low, self.actions_high, torch.clamp(current_action, -1.0, 1.0))
        return current_action

    def env_step(self, env, actions):
        """[ref: player.py:226-240] tensors stay on the device (`is_tensor_obses`)."""
        obs, rewards, dones, infos = env.step(actions)
        if self.value_size == 1 and rewards.dim() > 1:
            rewards = rewards.squeeze(-1)
        return obs["obs"], rewards, dones, infos

    def env_reset(self, env):
        return env.reset()["obs"]

    def reset(self) -> None:
        self.states = None

    # ---- evaluation loop  [ref: player.py:319-422] -----------------------------------------------------
    def run(self):
        """Plays until `games_num * n_game_life` episodes have finished; returns (av reward, av steps, games played).
        Accounting as in the reference: every env that finishes on a step is counted (so the last step can overshoot `games_num`),
        episode reward / length accumulate per env and are zeroed by the done flag."""
        n_games = self.games_num * self.n_game_life
        sum_rewards, sum_steps, games_played = 0.0, 0.0, 0
        for _ in range(n_games):
            if games_played >= n_games:
                break
            obses = self.env_reset(self.env)
            batch_size = int(obses["state"].shape[0]) if isinstance(obses, dict) else int(obses.shape[0])
            cr = torch.zeros(batch_size, dtype=torch.float32, device=self.device)
            steps = torch.zeros(batch_size, dtype=torch.float32, device=self.device)
            for _n in range(self.max_steps):
                action = self.get_action(obses, self.is_deterministic)
                obses, r, done, _info = self.env_step(self.env, action)
                cr += r.to(self.device)
                steps += 1
                fdone = done.to(self.device).float()
                fin = fdone[::self.num_agents]
                # one host read per step, as the reference's `done.nonzero()`: [count, reward sum, step sum] of the finished envs
                tally = torch.stack([fin.sum(), (cr[::self.num_agents] * fin).sum(), (steps[::self.num_agents] * fin).sum()]).tolist()
                done_count = int(tally[0])
                if done_count > 0:
                    games_played += done_count
                    cr = cr * (1.0 - fdone)
                    steps = steps * (1.0 - fdone)
                    sum_rewards += tally[1]
                    sum_steps += tally[2]
                    if self.print_stats:
                        print(f"reward: {tally[1] / done_count:.4} steps: {tally[2] / done_count:.4f}")
                    if batch_size // self.num_agents == 1 or games_played >= n_games:
                        break
        print(sum_rewards)
        av_reward = sum_rewards / games_played * self.n_game_life
        av_steps = sum_steps / games_played * self.n_game_life
        print("av reward:", av_reward, "av steps:", av_steps)
        return av_reward, av_steps, games_played

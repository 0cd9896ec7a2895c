"""CPU suite: the Tier-3 task oracle (GoToPose / KeepXY / TrackXYVelocity) vs goldens produced by the reference's own task classes
(oracle/make_golden.py:tier3)."""
import torch

from oracle import usv_oracle_t as TT

T = torch.from_numpy

SPECS = {
    "gotopose": TT.Tier3Config(task=TT.GO_TO_POSE, position_tolerance=0.5, kill_after_n_steps_in_tolerance=3, kill_dist=10.0),
    "gotopose_sq": TT.Tier3Config(task=TT.GO_TO_POSE, position_tolerance=0.5, kill_after_n_steps_in_tolerance=3, kill_dist=10.0, reward_mode=1,
                                  heading_reward_mode=0, position_scale=2.0, heading_scale=3.0, sig_gain=2.0),
    "keepxy": TT.Tier3Config(task=TT.KEEP_XY, position_tolerance=0.1, kill_after_n_steps_in_tolerance=500, kill_dist=8.0),
    "keepxy_lin": TT.Tier3Config(task=TT.KEEP_XY, position_tolerance=0.1, kill_after_n_steps_in_tolerance=1, kill_dist=8.0, reward_mode=0),
    "trackxyvel": TT.Tier3Config(task=TT.TRACK_XY_VELOCITY, lin_vel_tolerance=0.3, kill_after_n_steps_in_tolerance=2, kill_dist=9.0),
}


def test_tier3_tasks_vs_reference(golden):
    G = golden("tier3_tasks")
    for tag, c in SPECS.items():
        K, n = G[f"{tag}_pos"].shape[:2]
        goal = torch.zeros(n, dtype=torch.int32)
        prev_d = torch.zeros(n)
        th = T(G[f"{tag}_target_heading"]) if f"{tag}_target_heading" in G else torch.zeros(n)
        tv = T(G[f"{tag}_target_vel"]) if f"{tag}_target_vel" in G else torch.zeros((n, 2))
        for k in range(K):
            yaw = T(G[f"{tag}_yaw"][k])
            state = {"position": T(G[f"{tag}_pos"][k]), "orientation": torch.stack([torch.cos(yaw), torch.sin(yaw)], 1),
                     "linear_velocity": T(G[f"{tag}_vel"][k]), "angular_velocity": T(G[f"{tag}_w"][k])}
            obs, aux = TT.task_observation(c, state, T(G[f"{tag}_target"]), th, tv, T(G[f"{tag}_prev_action"][k]), T(G[f"{tag}_priv"][k]))
            assert torch.allclose(obs, T(G[f"{tag}_obs"][k]), rtol=1e-6, atol=1e-6), (tag, k)
            rew, die, prev_d = TT.task_reward_and_kills(c, aux, state, T(G[f"{tag}_actions"][k]), goal, prev_d)
            assert torch.allclose(rew, T(G[f"{tag}_reward"][k]), rtol=1e-5, atol=1e-6), (tag, k)
            assert torch.equal(die, T(G[f"{tag}_die"][k])) and torch.equal(goal, T(G[f"{tag}_goal_reached"][k])), (tag, k)
    # the crafted rows did their job
    assert G["gotopose_goal_reached"][-1][0] == 5 and G["gotopose_die"][0][1] == 1 and G["trackxyvel_die"][2][2] == 1

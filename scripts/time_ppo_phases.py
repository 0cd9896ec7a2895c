import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
from scripts.train_usv import make_env
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for tc in (True, False):
    env = make_env(UsvEnvConfig(num_envs=n).to_task_cfg(), "cuda:0", seed=1, collect_stats=False)
    env.env._task._nan_probe = False
    agent = A2CAgent(env, PPOConfig(seed=1), "cuda:0")
    agent.policy.tensor_cores = tc and agent.policy.tensor_cores
    for _ in range(4): agent.train_epoch()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tp = tu = 0.0
    for _ in range(10):
        with torch.no_grad():
            ev[0].record(); agent.play_steps(); ev[1].record(); agent._graph.replay(); ev[2].record()
        torch.cuda.synchronize()
        tp += ev[0].elapsed_time(ev[1]); tu += ev[1].elapsed_time(ev[2])
    # per-kernel timing of one minibatch step, eager
    pol, ds, mb = agent.policy, agent.ds, agent.minibatch_size
    s = slice(0, mb)
    def t(fn, reps=50):
        torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3
    g = lambda: pol.minibatch_grad(ds["obs"][s], ds["actions"][s], ds["old_logp_actions"][s], ds["advantages"][s], ds["old_values"][s], ds["returns"][s], ds["mu"][s], ds["sigma"][s])
    print(f"tensor_cores={tc}: play {tp/10:.2f} ms/epoch, update {tu/10:.2f} ms/epoch | minibatch_grad {t(g):.1f} us, optimizer_step(+pack) {t(pol.optimizer_step):.1f} us, "
          f"obs_rms.update {t(lambda: pol.obs_rms.update(ds['obs'][s])):.1f} us, act {t(lambda: pol.act(agent.obs)):.1f} us, env.step {t(lambda: env.env._task.engine.step(agent.buf['actions'][0])):.1f} us")

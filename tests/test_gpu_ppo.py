"""GPU parity of the PPO kernels (through the C ABI) against goldens produced by the reference's rl_games classes
and against the CPU oracle.  Bar: 1e-5 relative in fp32 (atol scaled to the quantity), written at each assert."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.rl.policy import PolicyMLP  # noqa: E402
from oracle import ppo_oracle as P  # noqa: E402
from tests.util import assert_close  # noqa: E402

DEV = "cuda:0"
D = 13
cu = lambda a: torch.as_tensor(a).to(DEV).contiguous()


def _policy(G, tensor_cores=False, **kw):
    """fp32 SIMT kernels by default: they carry the 1e-5 parity bar; the tensor-core path has its own tests below."""
    pol = PolicyMLP(D, DEV, lr=3e-4, tensor_cores=tensor_cores, **kw)
    pol.params.copy_(cu(G["params0"]))
    pol.obs_rms.load(G["obs_mean"], G["obs_var"], G["obs_count"])
    pol.val_rms.load(G["val_mean"], G["val_var"], G["val_count"])
    return pol


def test_param_contract(golden):
    G = golden("ppo")
    pol = PolicyMLP(D, DEV)
    assert pol.P == 18693 and PolicyMLP(33, DEV).P == 21253
    sd = pol.state_dict()
    assert [k for k in sd if k.startswith("a2c_network")] == [str(n) for n in G["param_names"]]
    assert sd["running_mean_std.running_mean_std.state.running_mean"].dtype == torch.float64
    v = pol.views()
    assert float(v["a2c_network.actor_mlp.0.bias"].abs().max()) == 0.0 and float(v["a2c_network.sigma"].abs().max()) == 0.0
    assert float(v["a2c_network.actor_mlp.2.weight"].abs().max()) <= 1 / np.sqrt(128) + 1e-7


def test_running_mean_std_update_vs_reference(golden):
    G = golden("ppo")
    pol = PolicyMLP(D, DEV)
    pol.obs_rms.update(cu(G["rms_update_obs"]))
    assert_close(pol.obs_rms.mean, G["obs_mean"], 1e-6, 1e-7); assert_close(pol.obs_rms.var, G["obs_var"], 1e-6, 1e-7)
    assert float(pol.obs_rms.count) == float(G["obs_count"])


def test_policy_inference_vs_reference(golden):
    G = golden("ppo")
    pol = _policy(G)
    M = G["obs"].shape[0]
    out = pol.act(cu(G["obs"]))
    assert_close(out["mus"], G["inf_mus"], 1e-5, 2e-6, "mus")
    assert_close(out["sigmas"], G["inf_sigmas"], 1e-6, 1e-7, "sigmas")
    assert_close(out["values"], G["inf_values"], 1e-5, 1e-5, "values (de-normalised)")
    # sampled actions follow a = mu + sigma * eps with the Philox/Box-Muller eps of the mirror; neglogp matches the
    # reference formula evaluated on the kernel's own actions
    eps = P.normal_eps(pol.seed, np.arange(M), 0)
    assert_close(out["actions"], T(G["inf_mus"]) + T(G["inf_sigmas"]) * eps, 1e-5, 1e-5, "actions")
    sg = out["sigmas"].cpu()
    nlp = P.neglogp(out["actions"].cpu(), out["mus"].cpu(), sg, torch.log(sg))
    assert_close(out["neglogpacs"], nlp, 1e-5, 1e-5, "neglogp")
    assert_close(pol.values(cu(G["obs"])), G["inf_values"], 1e-5, 1e-5, "get_values")
    # a second call draws fresh noise
    out2 = pol.act(cu(G["obs"]))
    assert not torch.equal(out2["actions"], out["actions"]) and torch.equal(out2["mus"], out["mus"])


T = torch.from_numpy


def test_minibatch_grad_adam_lr_vs_reference(golden):
    """3 consecutive PPO minibatch steps: loss terms, KL, full gradient, grad-norm, Adam update, adaptive lr."""
    G = golden("ppo")
    pol = _policy(G)
    obs, act = cu(G["obs"]), cu(G["mb_actions"])
    old_nlp, adv = cu(G["mb_old_neglogp"]), cu(G["mb_adv"])
    old_v, ret = cu(G["mb_old_values"]).reshape(-1).contiguous(), cu(G["mb_returns"]).reshape(-1).contiguous()
    mu, sigma = cu(G["mb_old_mu"]).clone(), cu(G["mb_old_sigma"]).clone()
    for it in range(3):
        g = pol.minibatch_grad(obs, act, old_nlp, adv, old_v, ret, mu, sigma)
        st = pol.stats()
        for k in ("a_loss", "c_loss", "entropy", "b_loss", "kl", "loss"):
            assert_close(torch.tensor(st[k]), G[f"it{it}_{k}"], 2e-5, 2e-6, f"it{it} {k}")
        # gradient entries span 1e-7..1e-1: 1e-5 relative + 2e-7 absolute (fp32 sums over 200 samples)
        assert_close(g[: pol.P], G[f"it{it}_grads"], 1e-4, 2e-7, f"it{it} grads")
        assert_close(mu, G[f"it{it}_mus"], 1e-5, 2e-6, "update_mu_sigma")
        pol.optimizer_step()
        st = pol.stats()
        assert_close(torch.tensor(st["grad_norm"]), G[f"it{it}_grad_norm"], 1e-5, 1e-7, "grad norm")
        assert abs(st["lr"] - float(G[f"it{it}_lr"])) < 1e-9
        assert_close(pol.params, G[f"it{it}_params_after"], 1e-5, 2e-7, f"it{it} params after Adam")
        assert abs(float(pol.lr) - float(G[f"it{it}_new_lr"])) < 1e-9
    assert int(pol.step) == 3


def test_minibatch_grad_vs_oracle_autograd_large():
    """8192-sample minibatch (the reference's minibatch_size) against torch autograd on the oracle."""
    torch.manual_seed(3)
    M = 8192 + 50
    pol = PolicyMLP(D, DEV, seed=5, tensor_cores=False)
    pol.params.add_(0.02 * torch.randn(pol.P, device=DEV))
    obs = torch.randn((M, D)) * 2
    pol.obs_rms.update(cu(obs[:500]))
    orc_rms = P.RunningMeanStd((D,)); orc_rms.mean, orc_rms.var = pol.obs_rms.mean.cpu(), pol.obs_rms.var.cpu()
    params = pol.params.cpu().clone().requires_grad_(True)
    inf = P.policy_inference(params.detach(), obs, D, orc_rms, P.RunningMeanStd((1,)), eps=torch.randn((M, 2)))
    batch = dict(obs=obs, actions=inf["actions"] + 0.2 * torch.randn((M, 2)), old_logp_actions=inf["neglogpacs"] + 0.1 * torch.randn(M),
                 advantages=torch.randn(M), old_values=torch.randn((M, 1)) * 0.3, returns=torch.randn((M, 1)) * 0.5,
                 mu=inf["mus"] + 0.02 * torch.randn((M, 2)), sigma=inf["sigmas"].clone())
    loss, st = P.minibatch_loss(params, batch, D, orc_rms)
    loss.backward()
    mu, sg = cu(batch["mu"]).clone(), cu(batch["sigma"]).clone()
    g = pol.minibatch_grad(cu(obs), cu(batch["actions"]), cu(batch["old_logp_actions"]), cu(batch["advantages"]),
                           cu(batch["old_values"]).reshape(-1).contiguous(), cu(batch["returns"]).reshape(-1).contiguous(), mu, sg)
    s = pol.stats()
    assert_close(torch.tensor(s["loss"]), loss.detach(), 2e-5, 2e-6, "loss")
    assert_close(torch.tensor(s["kl"]), st["kl"], 1e-4, 1e-6, "kl")
    assert_close(g[: pol.P], params.grad, 2e-4, 3e-7, "grads (fp32 sums over 8k samples)")
    assert_close(mu, st["mu"], 1e-5, 2e-6, "new mu")


def test_tensor_core_forward_vs_fp32(golden):
    """tcgen05/TMEM forward (TF32 operands, tanh.approx) against the fp32 SIMT kernel: ~1e-3 class agreement, identical
    sampling noise (same Philox keys), and against the reference golden at the same tolerance."""
    G = golden("ppo")
    tc, ref = _policy(G, tensor_cores=True), _policy(G, tensor_cores=False)
    assert tc.tensor_cores and not ref.tensor_cores
    for M in (200, 128, 1, 16384 + 77):
        obs = cu(G["obs"]) if M == 200 else torch.randn((M, D), device=DEV) * 3
        a, b = tc.act(obs), ref.act(obs)
        # hidden activations carry ~5e-4 relative error each; mu/value are O(0.1-1) sums of 128 terms
        assert_close(a["mus"], b["mus"], 5e-3, 5e-3, f"tc mus M={M}")
        assert_close(a["values"], b["values"], 5e-3, 1e-2, f"tc values M={M}")
        assert torch.equal(a["sigmas"], b["sigmas"])
        assert_close(a["actions"] - a["mus"], b["actions"] - b["mus"], 1e-6, 1e-6, "same noise")
        sg = a["sigmas"].cpu()
        assert_close(a["neglogpacs"], P.neglogp(a["actions"].cpu(), a["mus"].cpu(), sg, torch.log(sg)), 1e-5, 1e-5, "neglogp self-consistent")
        assert_close(tc.values(obs), b["values"], 5e-3, 1e-2, "tc get_values")
    out = tc.act(cu(G["obs"]))
    assert_close(out["mus"], G["inf_mus"], 5e-3, 5e-3, "tc mus vs reference")
    assert_close(out["values"], G["inf_values"], 5e-3, 1e-2, "tc values vs reference")


@pytest.mark.parametrize("D", [13, 33])
def test_tensor_core_minibatch_grad_vs_fp32(D):
    """tcgen05 training kernel (all GEMMs incl. the TMEM-resident weight-gradient accumulators) vs the fp32 SIMT kernel, for the
    classic (13 -> K = 16) and the live (33 -> K = 48, X / W1 tiles aliased with W2^T) observation widths."""
    torch.manual_seed(11)
    for M in (200, 8192, 8192 * 3 + 5, 128 * 200):      # ragged single tile / the reference minibatch / several tiles per CTA
        tc, ref = PolicyMLP(D, DEV, seed=5, tensor_cores=True), PolicyMLP(D, DEV, seed=5, tensor_cores=False)
        assert tc.tensor_cores
        delta = 0.02 * torch.randn(tc.P, device=DEV)
        tc.params.add_(delta); ref.params.add_(delta)
        obs = torch.randn((M, D), device=DEV) * 2
        for p in (tc, ref):
            p.obs_rms.update(obs[:150])
        inf = ref.act(obs)
        act = (inf["actions"] + 0.2 * torch.randn((M, 2), device=DEV)).contiguous()
        old_nlp = (inf["neglogpacs"] + 0.1 * torch.randn(M, device=DEV)).contiguous()
        adv, old_v, ret = torch.randn(M, device=DEV), torch.randn(M, device=DEV) * 0.3, torch.randn(M, device=DEV) * 0.5
        mu0 = (inf["mus"] + 0.02 * torch.randn((M, 2), device=DEV)).contiguous()
        mu_a, sg_a, mu_b, sg_b = mu0.clone(), inf["sigmas"].clone(), mu0.clone(), inf["sigmas"].clone()
        ga = tc.minibatch_grad(obs, act, old_nlp, adv, old_v, ret, mu_a, sg_a).clone()
        gb = ref.minibatch_grad(obs, act, old_nlp, adv, old_v, ret, mu_b, sg_b).clone()
        sa, sb = tc.stats(), ref.stats()
        for k in ("a_loss", "c_loss", "b_loss", "kl", "loss", "entropy"):
            assert abs(sa[k] - sb[k]) <= 5e-3 * abs(sb[k]) + 2e-4, (M, k, sa[k], sb[k])
        P_ = tc.P
        # TF32 + tanh.approx: compare per parameter block relative to that block's gradient scale
        off = 0
        for name, shp in zip(["sigma", "w1", "b1", "w2", "b2", "wv", "bv", "wmu", "bmu"], [(2,), (128, D), (128,), (128, 128), (128,), (1, 128), (1,), (2, 128), (2,)]):
            n = int(np.prod(shp))
            a, b = ga[off:off + n], gb[off:off + n]
            scale = float(b.abs().max()) + 1e-12
            err = float((a - b).abs().max())
            # 2% of the block's scale; + 5e-5 absolute for the scalar blocks (bias gradients are means of sign-alternating
            # per-sample terms, so TF32 noise does not shrink with the cancelling sum)
            assert err <= 2e-2 * scale + 5e-5, (M, name, err, scale)
            if n > 2:
                cos = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
                assert cos > 0.999, (M, name, cos)
            off += n
        assert_close(mu_a, mu_b, 5e-3, 5e-3, "new mu")
        tc.optimizer_step(); ref.optimizer_step()
        assert float((tc.params - ref.params).abs().max()) < 2.1e-4       # one Adam step moves each weight by <= lr


def test_ppo_loop_learns_on_tensor_cores():
    from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
    from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
    from scripts.train_usv import make_env
    cfg = UsvEnvConfig(num_envs=2048, max_episode_length=400)
    agent = A2CAgent(make_env(cfg.to_task_cfg(), DEV, seed=3, collect_stats=False), PPOConfig(seed=3), DEV)
    assert agent.policy.tensor_cores
    rewards = []
    for chunk in range(4):
        for _ in range(15):
            agent.train_epoch()
        rewards.append(agent.episode_stats()[0])
    assert torch.isfinite(agent.policy.params).all() and rewards[-1] > rewards[0], rewards


def test_peer_allreduce_world1_and_abi():
    """The NVLink peer all-reduce with a single rank degenerates to a copy through the IPC window: exercises window allocation,
    the flag protocol (own flag), sequence numbers across calls and graph capture.  Multi-rank: scripts/test_peer_allreduce.py."""
    from omniisaacgymenvs_loop_b200.rl.peer import PeerAllReduce
    ar = PeerAllReduce(5000, DEV, 0, 1)
    x = torch.randn(4097, device=DEV)
    for k in range(5):
        y = ar(x * (k + 1), torch.empty_like(x))
        assert torch.equal(y, x * (k + 1))
    g = torch.cuda.CUDAGraph()
    out = torch.empty_like(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(4):
            ar(x, out)
    g.replay(); g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, x) and int(ar.seq.item()) == 5 + 8
    ar.check()
    ar.close()


def test_fused_minibatch_step_equals_unfused():
    """ppo_minibatch_step_tc (T1 + T2 + one cooperative reduce / clip / Adam / lr / re-pack kernel) vs the unfused sequence
    minibatch_grad -> optimizer_step -> pack on the same tensor-core kernels, eager and inside a CUDA graph."""
    torch.manual_seed(3)
    for M in (8192, 1000):
        a, b = PolicyMLP(D, DEV, seed=5, tensor_cores=True), PolicyMLP(D, DEV, seed=5, tensor_cores=True)
        obs = torch.randn((M, D), device=DEV) * 2
        for p in (a, b):
            p.obs_rms.update(obs[:150])
        inf = a.act(obs)
        act = (inf["actions"] + 0.2 * torch.randn((M, 2), device=DEV)).contiguous()
        old_nlp = (inf["neglogpacs"] + 0.1 * torch.randn(M, device=DEV)).contiguous()
        adv, old_v, ret = torch.randn(M, device=DEV), torch.randn(M, device=DEV) * 0.3, torch.randn(M, device=DEV) * 0.5
        mu_a, sg_a, mu_b, sg_b = inf["mus"].clone(), inf["sigmas"].clone(), inf["mus"].clone(), inf["sigmas"].clone()
        for it in range(4):
            a.minibatch_grad(obs, act, old_nlp, adv, old_v, ret, mu_a, sg_a)
            a.optimizer_step()
            a.pack()
            b.minibatch_step(obs, act, old_nlp, adv, old_v, ret, mu_b, sg_b)
            # identical T1/T2 kernels; only the summation order of the gradient norm differs (clip coefficient ~1e-7 relative)
            assert_close(b.params, a.params, 1e-6, 1e-8, f"params M={M} it={it}")
            # after the first step the parameters differ in their last bits, and a last-bit difference that straddles a TF32 rounding
            # boundary of an operand is a 2^-11 relative change of that operand: gradients agree to ~1e-4 of their scale from then on
            gs = float(a.exp_avg.abs().max())
            assert_close(b.exp_avg, a.exp_avg, 1e-5 if it == 0 else 2e-4, 1e-10 if it == 0 else 2e-4 * gs, "exp_avg")
            assert_close(b.exp_avg_sq, a.exp_avg_sq, 1e-5 if it == 0 else 4e-4, 1e-14 if it == 0 else 4e-4 * gs * gs, "exp_avg_sq")
            assert_close(b.packed, a.packed, 1e-6, 1e-8, "packed tiles")
            sa, sb = a.stats(), b.stats()
            assert all(abs(sa[k] - sb[k]) <= 1e-5 * abs(sa[k]) + 1e-7 for k in sa), (sa, sb)
            assert int(a.step) == int(b.step) == it + 1
            # the fused tail sums the partial gradients in another order; from the second step on the parameters differ in their last
            # bits and TF32 rounding of the operands turns that into ~1e-5 of the mu scale
            assert_close(mu_b, mu_a, 1e-5, 1e-6 if it == 0 else 2e-5, "mu write-back")
    # capturable (cooperative launch inside a graph)
    g = torch.cuda.CUDAGraph()
    before = b.params.clone()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        b.minibatch_step(obs, act, old_nlp, adv, old_v, ret, mu_b, sg_b)
    g.replay(); g.replay()
    torch.cuda.synchronize()
    assert int(b.step) == 6 and not torch.equal(before, b.params) and torch.isfinite(b.params).all()


@pytest.mark.parametrize("world", [2, 4])
def test_fused_peer_minibatch_step_packet_protocol(world):
    """The multi-rank minibatch step (gradient all-reduce INSIDE the cooperative tail kernel, 8-byte {value, sequence} packets) on ONE
    GPU: the packets of the peer ranks are placed into rank 0's window beforehand (a cooperative kernel does not share the device
    with a second one, so in-process ranks cannot spin on each other), rank 0 runs its fused step and must (i) land on the unfused
    sequence  sum_r minibatch_grad_r -> Adam with 1/world  [ref: RLG/common/a2c_common.py:308-330], (ii) have stored its own packets,
    sequence-stamped, into every peer's window, (iii) alternate the slot parity and advance the sequence counter.  The concurrent
    NVLink run of the same kernel, with bit-identity across ranks: scripts/check_peer_allreduce.py (torchrun, >= 2 GPUs)."""
    from omniisaacgymenvs_loop_b200.rl.peer import InProcessStepExchange
    torch.manual_seed(11)
    M = 2048
    ex = InProcessStepExchange(D, DEV, world)
    ecap = int(ex.ranks[0].comm.cap) // (2 * world)
    pol = PolicyMLP(D, DEV, seed=5, tensor_cores=True, world_size=world)
    ref = PolicyMLP(D, DEV, seed=5, tensor_cores=True, world_size=world)
    P = pol.P
    data = []
    for r in range(world):
        obs = torch.randn((M, D), device=DEV) * 2
        inf = ref.act(obs)
        data.append(dict(obs=obs, act=(inf["actions"] + 0.2 * torch.randn((M, 2), device=DEV)).contiguous(),
                         nlp=(inf["neglogpacs"] + 0.1 * torch.randn(M, device=DEV)).contiguous(), adv=torch.randn(M, device=DEV),
                         old_v=torch.randn(M, device=DEV) * 0.3, ret=torch.randn(M, device=DEV) * 0.5,
                         mu=inf["mus"].clone(), sg=inf["sigmas"].clone(), mu_ref=inf["mus"].clone(), sg_ref=inf["sigmas"].clone()))
    step = lambda p, d, mu, sg, **kw: p(d["obs"], d["act"], d["nlp"], d["adv"], d["old_v"], d["ret"], d[mu], d[sg], **kw)
    win = [w.view(torch.int32) for w in ex.windows]
    for it in range(5):
        seq, par = it + 1, (it + 1) & 1
        local = [step(ref.minibatch_grad, data[r], "mu_ref", "sg_ref").clone() for r in range(world)]   # span a rank sends: grads + sums
        for r in range(1, world):                          # the peers' packets, as their tail kernels would have stored them
            base = (par * world + r) * ecap * 2
            win[0][base:base + 2 * (P + 5):2] = local[r][:P + 5].view(torch.int32)
            win[0][base + 1:base + 2 * (P + 5):2] = seq
        step(pol.minibatch_step, data[0], "mu", "sg", peer=ex.ranks[0])
        torch.cuda.synchronize()
        assert int(ex.ranks[0].err.item()) == 0 and int(ex.ranks[0].seq.item()) == seq
        for r in range(1, world):                          # rank 0's packets in the peers' windows, slot [parity][0]
            base = (par * world + 0) * ecap * 2
            got_v = win[r][base:base + 2 * (P + 5):2].view(torch.float32)
            got_s = win[r][base + 1:base + 2 * (P + 5):2]
            assert bool((got_s == seq).all()), "sequence stamp"
            gsc = float(local[0][:P].abs().max())
            # same partial sums as minibatch_grad's second stage, added in another (fixed) order
            assert_close(got_v, local[0][:P + 5], 1e-5 if it == 0 else 2e-4, 1e-6 * gsc if it == 0 else 2e-4 * gsc, "packets of rank 0")
            other = ((1 - par) * world + 0) * ecap * 2
            assert bool((win[r][other + 1:other + 2 * (P + 5):2] == (seq - 1 if seq > 1 else 0)).all()), "the other parity is untouched"
        ref.grads.copy_(torch.stack(local).sum(0))
        ref.optimizer_step()
        assert_close(pol.params, ref.params, 1e-6, 1e-8, f"params it={it}")
        gs = float(ref.exp_avg.abs().max())
        assert_close(pol.exp_avg, ref.exp_avg, 1e-5 if it == 0 else 2e-4, 1e-10 if it == 0 else 2e-4 * gs, "exp_avg")
        assert_close(pol.packed, ref.packed, 1e-6, 1e-8, "packed tiles")
        sa, sb = ref.stats(), pol.stats()
        assert all(abs(sa[k] - sb[k]) <= 1e-5 * abs(sa[k]) + 1e-7 for k in sa), (sa, sb)
        assert float(pol.lr) == float(ref.lr) and int(pol.step) == int(ref.step) == seq
        assert_close(data[0]["mu"], data[0]["mu_ref"], 1e-5, 1e-6 if it == 0 else 2e-5, "mu write-back")


def test_tensor_core_paths_at_live_obs_width():
    """D = 33 (the live task's observation): forward and the fused minibatch step on the K = 48 tensor-core kernels vs fp32."""
    Dw = 33
    torch.manual_seed(4)
    tc, ref = PolicyMLP(Dw, DEV, seed=9, tensor_cores=True), PolicyMLP(Dw, DEV, seed=9, tensor_cores=False)
    assert tc.tensor_cores and tc.P == 21253
    for M in (77, 16384, 128 * 300 + 5):
        obs = torch.randn((M, Dw), device=DEV) * 2
        for p in (tc, ref):
            p.obs_rms.update(obs[: min(M, 500)])
        a, b = tc.act(obs), ref.act(obs)
        assert_close(a["mus"], b["mus"], 5e-3, 5e-3, f"mus M={M}")
        assert_close(a["values"], b["values"], 5e-3, 1e-2, f"values M={M}")
        assert_close(a["actions"] - a["mus"], b["actions"] - b["mus"], 1e-6, 1e-6, "same noise")
    M = 8192
    obs = torch.randn((M, Dw), device=DEV) * 2
    inf = ref.act(obs)
    args = [obs, inf["actions"].contiguous(), (inf["neglogpacs"] + 0.1 * torch.randn(M, device=DEV)).contiguous(), torch.randn(M, device=DEV),
            torch.randn(M, device=DEV) * 0.3, torch.randn(M, device=DEV) * 0.5]
    for it in range(3):
        tc.minibatch_step(*args, inf["mus"].clone(), inf["sigmas"].clone())
        ref.minibatch_grad(*args, inf["mus"].clone(), inf["sigmas"].clone())
        ref.optimizer_step()
        sa, sb = tc.stats(), ref.stats()
        for k in ("a_loss", "c_loss", "b_loss", "kl", "loss"):
            assert abs(sa[k] - sb[k]) <= 1e-2 * abs(sb[k]) + 5e-4, (it, k, sa[k], sb[k])
        # each Adam step moves a weight by <= lr (sign flips of near-zero gradients: <= 2 lr apart), lr grows x1.5 per step at low KL
        assert float((tc.params - ref.params).abs().max()) < (it + 1) * 4e-4


def test_rollout_bookkeeping_kernel_vs_torch_ops():
    """ppo_rollout_bookkeep_f32 == the eager bookkeeping of play_steps (reward shaping, uint8 dones, running returns / lengths, epoch
    accumulators) and AverageMeter.update (the windowed mean pinned against rl_games' in test_ppo_oracle_cpu.py), over several steps."""
    import ctypes
    from omniisaacgymenvs_loop_b200 import _lib
    from omniisaacgymenvs_loop_b200.rl.a2c import AverageMeter
    n = 5003
    g = torch.Generator().manual_seed(3)
    cur_r, cur_l = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    acc = torch.zeros(3, dtype=torch.float64, device=DEV)
    mr, ml = AverageMeter(100, DEV), AverageMeter(100, DEV)
    t_r, t_l, t_acc = cur_r.clone(), cur_l.clone(), acc.clone()
    t_mr, t_ml = AverageMeter(100, DEV), AverageMeter(100, DEV)
    out_r, out_d = torch.empty(n, device=DEV), torch.empty(n, dtype=torch.uint8, device=DEV)
    for k in range(12):
        rew = torch.randn(n, generator=g).to(DEV)
        dones = (torch.rand(n, generator=g) < (0.0 if k == 0 else 0.02 * k)).long().to(DEV)      # first step: nothing finishes
        _lib.check(_lib.lib().ppo_rollout_bookkeep_f32(
            _lib.ptr(rew), _lib.ptr(dones), ctypes.c_float(0.01), _lib.ptr(out_r), _lib.ptr(out_d), _lib.ptr(cur_r), _lib.ptr(cur_l), _lib.ptr(acc),
            ctypes.c_void_p(mr.mean.data_ptr()), ctypes.c_void_p(mr.current_size.data_ptr()), ctypes.c_void_p(ml.mean.data_ptr()),
            ctypes.c_void_p(ml.current_size.data_ptr()), ctypes.c_float(100.0), ctypes.c_int64(n), _lib.stream()))
        # the eager formulation
        t_r += rew; t_l += 1
        d = dones.float()
        rs, ls, cnt = (t_r * d).sum(), (t_l * d).sum(), d.sum()
        t_acc += torch.stack([rs, ls, cnt]).double()
        t_mr.update(rs, cnt); t_ml.update(ls, cnt)
        t_r *= 1.0 - d; t_l *= 1.0 - d
        assert torch.equal(out_r, rew * 0.01) and torch.equal(out_d, dones.to(torch.uint8))
        assert torch.equal(cur_r, t_r) and torch.equal(cur_l, t_l)
        assert torch.allclose(acc, t_acc, rtol=1e-6, atol=1e-4) and float(acc[2]) == float(t_acc[2])
        assert float(mr.mean) == pytest.approx(float(t_mr.mean), rel=1e-5, abs=1e-5) and float(mr.current_size) == float(t_mr.current_size)
        assert float(ml.mean) == pytest.approx(float(t_ml.mean), rel=1e-5, abs=1e-5) and float(ml.current_size) == float(t_ml.current_size)
    assert float(mr.current_size) == 100.0 and float(acc[2]) > 500


class _StubVecEnv:
    """The two things A2CAgent.__init__ asks of an IVecEnv; update() never steps it."""

    def __init__(self, obs_dim, num_envs):
        import types
        self.env = types.SimpleNamespace(num_envs=num_envs)
        self._info = {"observation_space": {"state": types.SimpleNamespace(shape=(obs_dim,))}}

    def get_env_info(self):
        return self._info


@pytest.mark.parametrize("path", ["fp32", "fp32-graph", "tf32", "tf32-fused"])
def test_update_phase_vs_reference_train_epoch(golden, path):
    """Rows P2 / P3 / P5 end to end: A2CAgent.update() -- env-major dataset, value-normaliser schedule (values, then returns),
    advantage standardisation, obs normaliser updated in mini-epoch 0 only, in-order minibatches, mu / sigma write-back, per-minibatch
    adaptive-KL lr, Adam -- against two whole epochs of the reference's own ContinuousA2CBase.train_epoch / prepare_dataset / PPODataset
    on the same recorded rollouts (tests/golden/ppo_epoch.npz, oracle/make_golden.py:ppo_epoch)
    [ref: RLG/common/a2c_common.py:1197-1320, common/datasets.py:25-77].  fp32 kernels: 1e-5; the tcgen05 TF32 kernels at their own bar."""
    from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig, gae
    G = golden("ppo_epoch")
    Dd, Th, NA, MB, ME = (int(x) for x in G["shape"])
    tf32 = path.startswith("tf32")
    agent = A2CAgent(_StubVecEnv(Dd, NA), PPOConfig(horizon_length=Th, minibatch_size=MB, mini_epochs=ME, learning_rate=1e-4), DEV,
                     use_cuda_graph=False)
    pol = agent.policy
    pol.tensor_cores = pol.tensor_cores and tf32
    agent.fused_step = path == "tf32-fused"
    pol.params.copy_(cu(G["params0"])); pol._packed_dirty = True
    pol.obs_rms.load(G["obs_mean0"], G["obs_var0"], G["obs_count0"])
    pol.val_rms.load(G["val_mean0"], G["val_var0"], G["val_count0"])
    rt, at = (1e-5, 1e-6) if not tf32 else (2e-2, 2e-3)
    graph = None
    for ep in range(2):
        b = agent.buf
        for k, name in (("obses", "obses"), ("actions", "actions"), ("neglogpacs", "neglogpacs"), ("values", "values"), ("mus", "mus"),
                        ("sigmas", "sigmas"), ("dones", "dones")):
            b[k].copy_(cu(G[f"ep{ep}_{name}"]))
        b["rewards"].copy_(cu(G[f"ep{ep}_rewards"])[..., 0])
        # returns through the GAE kernel from the recorded rewards / values / dones: bit-exact vs the reference's discount_values
        gae(b["rewards"], b["values"].view(Th, -1), b["dones"], cu(G[f"ep{ep}_last_values"]).view(-1), cu(G[f"ep{ep}_last_dones"]),
            0.99, 0.95, agent.advs, agent.returns)
        assert torch.equal(agent.returns.cpu(), T(G[f"ep{ep}_returns"])[..., 0])
        if path == "fp32-graph":            # the whole update phase as one CUDA graph, as train_epoch replays it
            if graph is None:
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                snap = (pol.params.clone(), pol.exp_avg.clone(), pol.exp_avg_sq.clone(), pol._lr2.clone(), pol._step2.clone(),
                        [x.clone() for r in (pol.obs_rms, pol.val_rms) for x in (r.mean, r.var, r.count, r.mean32, r.var32)])
                agent.update()              # warm-up (allocations), then rewind and capture
                for dst, src in zip((pol.params, pol.exp_avg, pol.exp_avg_sq, pol._lr2, pol._step2), snap[:5]):
                    dst.copy_(src)
                for dst, src in zip([x for r in (pol.obs_rms, pol.val_rms) for x in (r.mean, r.var, r.count, r.mean32, r.var32)], snap[5]):
                    dst.copy_(src)
                for k in ("mus", "sigmas"):
                    b[k].copy_(cu(G[f"ep{ep}_{k}"]))
                with torch.cuda.graph(graph):
                    agent.update()
            graph.replay()
        else:
            agent.update()
        torch.cuda.synchronize()
        ds = agent.ds
        for k in ("obs", "actions", "old_logp_actions"):
            assert torch.equal(ds[k].cpu(), T(G[f"ep{ep}_ds_{k}"])), k                     # env-major flatten: bit-exact
        assert_close(ds["old_values"], G[f"ep{ep}_ds_old_values"][:, 0], 1e-5, 1e-6, "normalised old values")
        assert_close(ds["returns"], G[f"ep{ep}_ds_returns"][:, 0], 1e-5, 1e-6, "normalised returns")
        assert_close(ds["advantages"], G[f"ep{ep}_ds_advantages"], 1e-5, 2e-6, "standardised advantages")
        assert_close(pol.val_rms.mean, G[f"ep{ep}_val_mean"], 1e-6, 1e-7); assert_close(pol.val_rms.var, G[f"ep{ep}_val_var"], 1e-6, 1e-7)
        assert float(pol.val_rms.count) == float(G[f"ep{ep}_val_count"])
        assert_close(pol.obs_rms.mean, G[f"ep{ep}_obs_mean"], 1e-6, 1e-7); assert_close(pol.obs_rms.var, G[f"ep{ep}_obs_var"], 1e-6, 1e-7)
        assert float(pol.obs_rms.count) == float(G[f"ep{ep}_obs_count"])                     # + one minibatch per minibatch of mini-epoch 0
        assert int(pol.step) == (ep + 1) * ME * (Th * NA // MB)
        assert abs(float(pol.lr) - float(G[f"ep{ep}_last_lr"])) <= 1e-6 * float(pol.lr), "adaptive-KL schedule"
        if not tf32:
            assert_close(pol.params, G[f"ep{ep}_params"], rt, at, f"parameters after epoch {ep}")
        else:
            # Adam's m / sqrt(v) is sign-like where a gradient is small against its TF32 rounding noise, so single parameters may move
            # by up to lr per step the other way; the bar: the parameter UPDATE of the epoch points the reference's way (cosine), 99.5 %
            # of the parameters within 2 % / 2e-4, and nothing further off than the steps taken allow
            want, got, prev = T(G[f"ep{ep}_params"]).double(), pol.params.cpu().double(), T(G["params0" if ep == 0 else "ep0_params"]).double()
            dw, dg = want - prev, got - (prev if ep == 0 else prev0_got)
            cos = float((dw * dg).sum() / (dw.norm() * dg.norm()))
            assert cos > 0.98, f"update direction, epoch {ep}: cosine {cos}"
            bad = (got - want).abs() > 2e-4 + 2e-2 * want.abs()
            assert float(bad.double().mean()) < 5e-3, f"{int(bad.sum())} parameters outside 2e-2 / 2e-4"
            assert float((got - want).abs().max()) < (ep + 1) * ME * (Th * NA // MB) * 2.5e-4, "further than the Adam steps allow"
        prev0_got = pol.params.cpu().double()
        assert_close(ds["mu"], G[f"ep{ep}_ds_mu_final"], rt, 1e-5 if not tf32 else 5e-3, "mu write-back")
        assert_close(ds["sigma"], G[f"ep{ep}_ds_sigma_final"], rt, at, "sigma write-back")
        if not tf32:
            assert_close(pol.exp_avg, G[f"ep{ep}_exp_avg"], 1e-4, 1e-7, "exp_avg"); assert_close(pol.exp_avg_sq, G[f"ep{ep}_exp_avg_sq"], 1e-4, 1e-10)

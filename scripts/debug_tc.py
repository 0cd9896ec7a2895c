import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from omniisaacgymenvs_loop_b200.rl.policy import PolicyMLP
DEV="cuda:0"; D=13
torch.manual_seed(11)
M = int(sys.argv[1]) if len(sys.argv)>1 else 200
tc, ref = PolicyMLP(D, DEV, seed=5, tensor_cores=True), PolicyMLP(D, DEV, seed=5, tensor_cores=False)
delta = 0.02 * torch.randn(tc.P, device=DEV)
tc.params.add_(delta); ref.params.add_(delta)
obs = torch.randn((M, D), device=DEV) * 2
for p in (tc, ref): p.obs_rms.update(obs[:150])
inf = ref.act(obs)
act = (inf["actions"] + 0.2 * torch.randn((M, 2), device=DEV)).contiguous()
old_nlp = (inf["neglogpacs"] + 0.1 * torch.randn(M, device=DEV)).contiguous()
adv, old_v, ret = torch.randn(M, device=DEV), torch.randn(M, device=DEV) * 0.3, torch.randn(M, device=DEV) * 0.5
mu0 = (inf["mus"] + 0.02 * torch.randn((M, 2), device=DEV)).contiguous()
mu_a, sg_a, mu_b, sg_b = mu0.clone(), inf["sigmas"].clone(), mu0.clone(), inf["sigmas"].clone()
ga = tc.minibatch_grad(obs, act, old_nlp, adv, old_v, ret, mu_a, sg_a).clone()
gb = ref.minibatch_grad(obs, act, old_nlp, adv, old_v, ret, mu_b, sg_b).clone()
off=0
for name, shp in zip(["sigma", "w1", "b1", "w2", "b2", "wv", "bv", "wmu", "bmu"], [(2,), (128, D), (128,), (128, 128), (128,), (1, 128), (1,), (2, 128), (2,)]):
    n = int(np.prod(shp)); a, b = ga[off:off+n], gb[off:off+n]
    print(f"{name:6s} scale {float(b.abs().max()):.3e} err {float((a-b).abs().max()):.3e} |a|max {float(a.abs().max()):.3e} cos {float(torch.nn.functional.cosine_similarity(a,b,dim=0)):.5f}")
    if name in ("w1","w2"):
        A=a.view(shp); B=b.view(shp)
        print("   tc  row0:", A[0,:6].tolist()); print("   ref row0:", B[0,:6].tolist())
        # check transposed / permuted hypotheses
        if name=="w2": print("   cos with transpose:", float(torch.nn.functional.cosine_similarity(A.t().reshape(-1), B.reshape(-1), dim=0)))
    off+=n

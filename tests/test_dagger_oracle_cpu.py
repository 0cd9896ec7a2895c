"""The DAgger / SysID distillation oracle (oracle/dagger_oracle.py) against goldens produced by the reference's own StateHistoryEncoder,
USVSysIDAgent and USVSysIDTrainer (oracle/make_golden_dagger.py).  The parity gate of the CUDA path (csrc/dagger_sysid.cu, tests/test_gpu_dagger.py)."""
import os

import numpy as np
import torch

from oracle import dagger_oracle as DO

G = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "dagger_sysid.npz")).items()}
OD, LAT = 25, 8


def unflat(flat, shapes):
    out, off = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(flat[off:off + n].reshape(s).clone())
        off += n
    assert off == flat.numel()
    return out


def test_history_encoder_forward_matches_reference():
    for T in (50, 20, 10):
        p = unflat(G[f"enc{T}_params"], DO.history_encoder_shapes(OD, T, LAT))
        got = DO.history_encoder_forward(p, G[f"enc{T}_in"], T)
        assert torch.allclose(got, G[f"enc{T}_out"], rtol=1e-6, atol=1e-7), T
    # the reference reshapes (bs*T, 32) -> (bs, 32, T) without transposing: a transposed variant must NOT reproduce the golden
    p = unflat(G["enc50_params"], DO.history_encoder_shapes(OD, 50, LAT))
    x = torch.nn.functional.leaky_relu(torch.nn.functional.linear(G["enc50_in"].reshape(6 * 50, -1), p[0], p[1]))
    assert not torch.allclose(x.reshape(6, 50, 32).transpose(1, 2), x.reshape(6, 32, 50))


def _mlp(flat, dims, last):
    shapes = [s for a, b in zip(dims[:-1], dims[1:]) for s in ((b, a), (b,))]
    p = unflat(flat, shapes)

    def f(x):
        for i in range(0, len(p), 2):
            x = torch.nn.functional.linear(x, p[i], p[i + 1])
            x = last(x) if i == len(p) - 2 else torch.nn.functional.leaky_relu(x)
        return x
    return f


def test_student_action_and_teacher_latent_match_reference():
    p = unflat(G["tr_params0"], DO.history_encoder_shapes(OD, 50, LAT))
    head = _mlp(G["head_params"], [OD + LAT, 128, 128, 2], torch.tanh)
    got = DO.student_action(p, head, G["tr_sysid_obs"][0], 50, OD)
    assert torch.allclose(got, G["tr_actions0"], rtol=1e-6, atol=1e-7)
    teacher = _mlp(G["teacher_params"], [8, 64, 16, LAT], torch.nn.functional.leaky_relu)
    assert torch.allclose(teacher(G["tr_priv"].reshape(-1, 8)).reshape(G["tr_zstar"].shape), G["tr_zstar"], rtol=1e-6, atol=1e-7)


def test_trainer_updates_match_reference():
    """Two USVSysIDTrainer.update() calls: 4 epochs x 4 in-order minibatches of MSE + Adam each; parameters and diagnostics."""
    p = unflat(G["tr_params0"], DO.history_encoder_shapes(OD, 50, LAT))
    tr = DO.SysIDTrainerOracle(p, 50, LAT)
    hist = G["tr_sysid_obs"][:, :, : 50 * OD]
    for upd in range(2):
        m = tr.update(hist, G["tr_zstar"])
        flat = torch.cat([q.detach().reshape(-1) for q in tr.p])
        want = G[f"tr_params{upd + 1}"]
        assert float((flat - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max())), upd
        row = G["tr_metrics"][upd]
        got = [m["mse"], m["zstar_var_mean"], m["zhat_var_mean"], m["r2_total"]] + [m[f"r2_dim{i}"] for i in range(LAT)]
        assert np.allclose(np.asarray(got), row.numpy(), rtol=2e-4, atol=1e-5), (upd, got, row)
    assert tr.lr == 5e-4 and abs(float(G["tr_lr_after"]) - 5e-4) < 1e-9      # StepLR(200, 0.1): unchanged after two updates
    tr.itr = 200
    assert abs(tr.lr - 5e-5) < 1e-12

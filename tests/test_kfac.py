"""KFAC curvature statistics (second half).  CPU: the oracle's restatement of what the reference's
estimator extracts (tracer.py:196-332, curvature_blocks.py:262-281, curvature_tags_and_blocks.py:142-156) is checked
against the parameter gradients of the same oracle (sum_rows x^T dy is the weight gradient), and the block algebra of
deepsolid_b200.kfac (moving average, pi-adjusted inverse) against closed forms.  GPU: ds_kfac_factors through the
C ABI against the oracle on the same walkers and parameters, 1e-9 relative to the largest entry of each factor."""
import math

import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C
from deepsolid_b200 import kfac
from oracle import deepsolid_oracle as O


def test_oracle_factors_are_consistent_with_parameter_gradients():
    sc, kl, pn, P = system("h4")
    X = torch.as_tensor(C.init_walkers(sc, 3, seed=11))
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    F = O.kfac_factors(f, P, X)
    ones, zeros = torch.ones(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64)
    g_abs = O.logpsi_vjp(f, P, X, ones, zeros)
    g_ph = O.logpsi_vjp(f, P, X, zeros, ones)
    s2 = math.sqrt(2.0)
    n_e = sum(sc.nelec)
    for kind, rows_per_walker in (("single", n_e), ("double", n_e * n_e), ("orbital", None)):
        for i, blk in enumerate(F[kind]):
            a, g = blk["inputs_factor"], blk["outputs_factor"]
            assert torch.allclose(a, a.T) and torch.allclose(g, g.T)
            assert float(torch.linalg.eigvalsh(a).min()) > -1e-12 and float(torch.linalg.eigvalsh(g).min()) > -1e-12
            has_b = "b" in P[kind][i]
            if has_b:
                assert abs(float(a[-1, -1]) - 1.0) < 1e-14          # E[1 * 1] of the homogeneous coordinate
            if rows_per_walker is not None:
                assert blk["extra_scale"] == rows_per_walker
            else:
                assert blk["extra_scale"] == sc.nelec[i]
            # sum_rows (x,1)^T dy = sqrt2 (dRe F - i d angle) / d(w; b),  F = conj(log psi)
            wb_abs = torch.cat([g_abs[kind][i]["w"], g_abs[kind][i]["b"][None]]) if has_b else g_abs[kind][i]["w"]
            wb_ph = torch.cat([g_ph[kind][i]["w"], g_ph[kind][i]["b"][None]]) if has_b else g_ph[kind][i]["w"]
            want = s2 * torch.complex(wb_abs, -wb_ph)
            assert float((blk["xdy"] - want).abs().max()) < 1e-10 * max(1.0, float(want.abs().max()))
    for s, env in enumerate(F["envelope"]):
        for leaf in ("pi", "sigma"):
            dw = s2 * torch.complex(g_abs["envelope"][s][leaf], -g_ph["envelope"][s][leaf])
            assert float((env[leaf] - dw * dw / 3).abs().max()) < 1e-12 * max(1.0, float(dw.abs().max()) ** 2)


def test_weighted_moving_average_and_pi_adjusted_inverse():
    torch.manual_seed(0)
    m = kfac.WeightedMovingAverage(torch.zeros(2, 2, dtype=torch.float64))
    a, b = torch.eye(2, dtype=torch.float64), 3.0 * torch.eye(2, dtype=torch.float64)
    m.update(a, 0.95, 1.0)
    assert torch.allclose(m.value, a)                                # the first estimate is unbiased
    m.update(b, 0.95, 1.0)
    assert torch.allclose(m.value, (0.95 * a + b) / 1.95)
    # (A kron G + damping I)^-1 is approximated by (A + pi d I)^-1 kron (G + d/pi I)^-1 with pi^2 = (tr A/n)/(tr G/m)
    A = torch.randn(5, 5, dtype=torch.float64); A = A @ A.T
    G = torch.randn(3, 3, dtype=torch.float64); G = G @ G.T
    damping = 1e-2
    ai, gi = kfac.pi_adjusted_inverse(A, G, damping)
    pi = math.sqrt((float(torch.trace(A)) / 5) / (float(torch.trace(G)) / 3))
    want_a = torch.linalg.inv(A + pi * math.sqrt(damping) * torch.eye(5, dtype=torch.float64))
    want_g = torch.linalg.inv(G + math.sqrt(damping) / pi * torch.eye(3, dtype=torch.float64))
    assert torch.allclose(torch.kron(ai, gi), torch.kron(want_a, want_g), rtol=1e-9, atol=1e-12)
    z0, z1 = kfac.pi_adjusted_inverse(torch.zeros(2, 2, dtype=torch.float64), G, damping)
    assert torch.allclose(z0, torch.eye(2, dtype=torch.float64) / math.sqrt(damping))


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("name,nw,opts", [
    ("h4", 5, {}),
    ("lih_prim", 4, {}),
    ("graphene8", 3, {}),
    ("h4", 3, {"bias_orbitals": True}),
    ("h4", 3, {"distance_type": "tri"}),
    ("h4", 3, {"full_det": True}),
    ("graphite54", 2, {}),
    ("h4", 3, {"use_last_layer": True}),
    ("graphene8", 3, {"use_last_layer": True, "bias_orbitals": True}),
])
def test_gpu_kfac_factors_match_oracle(name, nw, opts):
    from deepsolid_b200 import network
    sc, kl, pn, P = system(name)
    if opts:
        pn2 = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec, **opts)
        P = O.params_to_torch(pn2)
    net = network.make_solid_fermi_net(envelope_type="isotropic", klist=kl, simulation_cell=sc, determinants=8,
                                       method_name="eval_logdet", full_det=opts.get("full_det", False),
                                       bias_orbitals=opts.get("bias_orbitals", False),
                                       distance_type=opts.get("distance_type", "nu"),
                                       use_last_layer=opts.get("use_last_layer", False))
    hp = net.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=21))
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet", full_det=opts.get("full_det", False),
                               bias_orbitals=opts.get("bias_orbitals", False),
                               distance_type=opts.get("distance_type", "nu"))
    want = O.kfac_factors(f, P, X)
    got = kfac.curvature_estimate(hp, P, X.cuda(), sync=False)
    worst = 0.0
    for kind in ("single", "double", "orbital"):
        assert len(got[kind]) == len(want[kind])
        for gb, wb in zip(got[kind], want[kind]):
            assert gb["extra_scale"] == wb["extra_scale"]
            for key in ("inputs_factor", "outputs_factor"):
                assert gb[key].shape == wb[key].shape, (kind, key, gb[key].shape, wb[key].shape)
                worst = max(worst, _rel(gb[key].cpu(), wb[key]))
    for ge, we in zip(got["envelope"], want["envelope"]):
        for leaf in ("pi", "sigma"):
            worst = max(worst, _rel(ge[leaf].cpu(), we[leaf]))
    assert worst < 1e-9, worst


@pytest.mark.gpu
def test_gpu_kfac_factors_chunk_invariance_and_empty():
    from deepsolid_b200 import network
    sc, kl, pn, P = system("h4")
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, 37, seed=5)).cuda()
    a = kfac.curvature_estimate(hp, P, X, sync=False)
    hp.set_workspace_limit(2 << 20)                      # a few walkers per chunk
    b = kfac.curvature_estimate(hp, P, X, sync=False)
    hp.set_workspace_limit(24 << 30)
    for kind in ("single", "double", "orbital"):
        for x, y in zip(a[kind], b[kind]):
            assert _rel(x["inputs_factor"], y["inputs_factor"]) < 1e-12
            assert _rel(x["outputs_factor"], y["outputs_factor"]) < 1e-12
    raw = hp.kfac_factors(X[:0])
    assert raw["batch"] == 0 and float(raw["single"][0]["a"].abs().max()) == 0.0


@pytest.mark.gpu
def test_gpu_kfac_step_follows_the_block_algebra_on_oracle_factors():
    """One optimiser step on the device equals the same update computed from the oracle's factors and gradient."""
    from deepsolid_b200 import network, train
    sc, kl, pn, P = system("h4")
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, 8, seed=3))
    loss_fn = train.make_loss(net.apply, None, sc, clip_local_energy=5.0, clip_type="real", mode="for")
    opt = kfac.Optimizer(loss_fn.value_and_grad, hp, norm_constraint=1e-3)
    new_params, stats = opt.step(P, X.cuda(), learning_rate=5e-2, damping=1e-3)
    assert 0.0 < stats["coefficient"] <= 1.0
    # the same step from oracle-side statistics
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    el = O.local_energy_seperate(O.make_solid_fermi_net(kl, sc, method_name="eval_logdet"), sc, mode="for")
    _, _, g = O.total_energy_value_and_grad(f, el, P, X, clip_local_energy=5.0, clip_type="real")
    F = O.kfac_factors(f, P, X)
    lr, damping = 5e-2, 1e-3
    pre, leaves_g = [], []
    for kind in ("single", "double", "orbital"):
        for blk, gl in zip(F[kind], g[kind]):
            has_b = "b" in gl
            ai, gi = kfac.pi_adjusted_inverse(blk["inputs_factor"], blk["outputs_factor"], damping / blk["extra_scale"])
            vec = torch.cat([gl["w"], gl["b"][None]]) if has_b else gl["w"]
            res = ai @ vec @ gi / blk["extra_scale"]
            pre += [res[:-1], res[-1]] if has_b else [res]
            leaves_g += [gl["w"], gl["b"]] if has_b else [gl["w"]]
    for env, gl in zip(F["envelope"], g["envelope"]):
        for leaf in ("pi", "sigma"):
            pre.append((gl[leaf] / (env[leaf] + damping)).real)
            leaves_g.append(gl[leaf])
    sq = sum(float((a * b).sum()) for a, b in zip(pre, leaves_g)) * lr ** 2
    coeff = min(math.sqrt(1e-3 / sq), 1.0)
    assert abs(coeff - stats["coefficient"]) < 1e-6 * coeff
    want = [p - lr * coeff * d for p, d in zip(O._leaves(P), pre)]
    got = O._leaves({k: [{kk: vv.cpu() for kk, vv in d.items()} for d in new_params[k]] for k in new_params})
    for a, b in zip(got, want):
        assert _rel(a, b) < 1e-7


@pytest.mark.gpu
def test_gpu_kfac_momentum_follows_the_reference_recursion():
    """optimizer.py:444-463, 617-627: delta = -lr c F^-1 g + momentum * previous delta; velocities <- delta; the first
    step treats the momentum as 0.  Two optimisers (momentum 0.5 / 0) share step 1, so their curvature states agree and
    step 2 differs by exactly 0.5 * delta_1."""
    from deepsolid_b200 import network, train
    sc, kl, pn, P = system("h4")
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, 8, seed=3)).cuda()
    loss_fn = train.make_loss(net.apply, None, sc, clip_local_energy=5.0, clip_type="real", mode="for")
    flat = lambda p: [t.detach().cpu().double() for t in kfac.flatten_params(p)]
    outs = {}
    for m in (0.5, 0.0):
        opt = kfac.Optimizer(loss_fn.value_and_grad, hp, norm_constraint=1e-3)
        p1, s1 = opt.step(P, X, learning_rate=5e-2, damping=1e-3, momentum=m)
        assert s1["momentum"] == 0.0
        p2, s2 = opt.step(p1, X, learning_rate=5e-2, damping=1e-3, momentum=m)
        assert s2["momentum"] == m
        outs[m] = (flat(p1), flat(p2))
    p0 = [torch.as_tensor(t).double() for t in kfac.flatten_params(P)]
    for a0, a1, b1, a2, b2 in zip(p0, outs[0.5][0], outs[0.0][0], outs[0.5][1], outs[0.0][1]):
        # (the envelope gradients are accumulated with atomics: equal to rounding, not bit for bit)
        assert float((a1 - b1).abs().max()) <= 1e-12 * max(float(a1.abs().max()), 1.0)
        delta1 = a1 - a0.cpu()
        err = float((a2 - b2 - 0.5 * delta1).abs().max())
        assert err <= 1e-6 * float(delta1.abs().max()) + 1e-13, (err, float(delta1.abs().max()))

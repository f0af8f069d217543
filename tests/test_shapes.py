"""Network shapes other than the benchmark's ((256,32),)*3 / 8 determinants (network.py:60-186 takes any
hidden_dims / determinants): 2 and 4 layers, narrower streams, 16 and 2 determinants, on the same kernels.  The CUDA
path needs equal widths per layer, hidden_two <= 32 and at most 4 layers; everything else must match the oracle with
the default tolerances (log|psi|, phase 1e-10; E_L 1e-8 Ha; gradients and KFAC factors 1e-9 relative)."""
import numpy as np
import pytest
import torch

from conftest import angle_diff
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O

SHAPES = [
    ("h4", ((128, 16),) * 2, 4),
    ("graphene8", ((256, 32),) * 4, 2),
    ("h4", ((64, 8),) * 3, 16),
    ("lih_prim", ((96, 12),) * 2, 8),          # widths the int8 tile shapes do not cover: fp64 DMMA kernels
]


def test_unsupported_shapes_raise():
    from deepsolid_b200 import network
    sc = C.build_system("h4")
    kl = C.make_klist(sc)
    for hd in [((256, 32), (128, 32)), ((256, 64),) * 2, ((256, 32),) * 5, ((256, 32),), ((255, 32),) * 2]:
        with pytest.raises(ValueError):
            network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                         determinants=8, hidden_dims=hd, method_name="eval_logdet")


@pytest.mark.gpu
@pytest.mark.parametrize("name,hidden,ndet", SHAPES)
def test_gpu_other_shapes_match_oracle(name, hidden, ndet):
    from deepsolid_b200 import network, hamiltonian, kfac
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    pn = O.init_params(np.random.default_rng(41), sc.original_cell.natm, sc.nelec, hidden_dims=hidden, determinants=ndet)
    P = O.params_to_torch(pn)
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=ndet, hidden_dims=hidden)
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    okw = dict(hidden_dims=hidden, determinants=ndet)
    f_ld = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet", **okw)
    f_ps = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet", **okw)
    nw = 3
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=23))
    v = ld.apply(P, X.to(dev)).cpu()
    vo = torch.stack([f_ld(P, x) for x in X])
    assert float((v.real - vo.real).abs().max()) < 1e-10
    assert float(angle_diff(v.imag, vo.imag).max()) < 1e-10
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        ko, eo = elo(P, X[b])
        assert abs(complex(ke[b].cpu()) - complex(ko)) < 1e-8 and abs(float(ew[b]) - float(eo)) < 1e-8
    # parameter gradient
    ca = torch.tensor([0.3, -1.1, 0.8], dtype=torch.float64)
    cp = torch.tensor([-0.6, 0.2, 0.5], dtype=torch.float64)
    hp.set_params(P)
    got = hp.logpsi_vjp(X.to(dev), ca.to(dev), cp.to(dev))
    want = O.logpsi_vjp(f_ps, P, X, ca, cp)
    for a, b in zip(O._leaves(got), O._leaves(want)):
        assert float((a.cpu() - b).abs().max()) <= 1e-9 * max(float(b.abs().max()), 1e-30)
    # KFAC statistics
    gf = kfac.curvature_estimate(hp, P, X.to(dev), sync=False)
    wf = O.kfac_factors(f_ps, P, X)
    for kind in ("single", "double", "orbital"):
        assert len(gf[kind]) == len(wf[kind])
        for g, w in zip(gf[kind], wf[kind]):
            for key in ("inputs_factor", "outputs_factor"):
                assert float((g[key].cpu() - w[key]).abs().max()) <= 1e-9 * max(float(w[key].abs().max()), 1e-30), (kind, key)

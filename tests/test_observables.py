"""Observables of estimator.py:15-85 (complex polarisation, structure factor).  CPU: properties of the oracle's
restatement.  GPU: ds_rho_q through the C ABI + the host means against the oracle, 1e-12."""
import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O


def test_oracle_observable_properties():
    sc, kl, pn, P = system("h4")
    X = torch.as_tensor(C.init_walkers(sc, 16, seed=2))
    sk = O.make_structure_factor(sc, nq=3)(X)
    assert sk.shape == (27,)
    assert abs(float(sk[0])) < 1e-12                     # q = 0: rho = N_e for every walker, no fluctuation
    assert float(sk.min()) > -1e-12
    # translating every electron by a lattice vector changes neither observable
    shift = torch.as_tensor(np.tile(sc.a[0], sum(sc.nelec)))
    assert torch.allclose(O.make_structure_factor(sc, nq=3)(X + shift), sk, atol=1e-10)
    pol = O.make_complex_polarization(sc, direction=0)
    assert abs(complex(pol(X + shift)) - complex(pol(X))) < 1e-10
    assert abs(complex(pol(X))) <= 1.0 + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8", "graphite54"])
def test_gpu_observables_match_oracle(name):
    from deepsolid_b200 import estimator, network
    sc, kl, pn, P = system(name)
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, 33, seed=4))
    sk = estimator.make_structure_factor(sc, nq=4, hotpath=hp)(X.cuda()).cpu()
    want = O.make_structure_factor(sc, nq=4)(X)
    assert float((sk - want).abs().max()) < 1e-12 * max(1.0, float(want.abs().max()))
    for direction in range(3):
        pol = estimator.make_complex_polarization(sc, direction=direction, hotpath=hp)(X.cuda()).cpu()
        assert abs(complex(pol) - complex(O.make_complex_polarization(sc, direction)(X))) < 1e-12
    # host walkers and the empty batch
    assert estimator.make_structure_factor(sc, nq=2, hotpath=hp)(X).shape == (8,)
    assert hp.rho_q(X[:0].cuda(), np.zeros((2, 3))).shape == (0, 2)

"""End-to-end use of the closures the way the reference driver uses them (process.py:249-372): Metropolis sweep ->
energy gradient -> KFAC step, repeated.  The variational energy of a randomly initialised H4 chain must go down."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


@pytest.mark.gpu
def test_vmc_loop_lowers_the_energy(tmp_path):
    import train_vmc
    from deepsolid_b200 import checkpoint
    hist, _ = train_vmc.run(system="h4", batch=256, iterations=40, burn_in=10, log=None, ckpt_dir=str(tmp_path),
                            structure_factor=True)
    e = np.array([h["energy"] for h in hist])
    assert np.all(np.isfinite(e)) and all(np.isfinite(h["variance"]) for h in hist)
    assert all(0.0 < h["pmove"] <= 1.0 for h in hist)
    assert e[-8:].mean() < e[:8].mean() - 3.0 * e[:8].std() / np.sqrt(8), (e[:8], e[-8:])
    assert hist[-1]["structure_factor"].shape == (8,)
    last = checkpoint.find_last_checkpoint(str(tmp_path))
    assert last is not None and last.endswith("qmcjax_ckpt_000040.npz")


@pytest.mark.gpu
def test_adam_training_step_lowers_the_energy():
    """The 'adam' branch of the reference driver (process.py:205-246) through train.make_training_step."""
    import torch
    from deepsolid_b200 import cell as C, network, qmc, train
    sc = C.build_system("h4")
    kl = C.make_klist(sc)
    P = network.init_solid_fermi_net_params(888, atoms=sc.original_cell.atom_coords(), spins=sc.nelec)
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    logdet = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = logdet.apply.hotpath()
    slog = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
    P = {k: [{kk: torch.as_tensor(v).to(hp.tdev) for kk, v in d.items()} for d in P[k]] for k in P}
    data = torch.as_tensor(C.init_walkers(sc, 256, seed=666)).to(hp.tdev)
    mcmc_step = qmc.make_mcmc_step(slog.apply, 256, sc.a, steps=10)
    loss_fn = train.make_loss(logdet.apply, None, sc, clip_local_energy=5.0, clip_type="real", mode="for")
    init, opt_update = train.make_adam_update(train.learning_rate_schedule(rate=1e-3))
    step = train.make_training_step(mcmc_step, loss_fn.value_and_grad, opt_update)
    state = init(P)
    for t in range(10):
        data, _ = mcmc_step(P, data, 100 + t, 0.02)
    e = []
    for t in range(40):
        data, P, state, loss, aux, pmove, sd = step(t, data, P, state, 200 + t, 0.02)
        e.append(float(loss))
    e = np.array(e)
    assert np.all(np.isfinite(e)) and e[-8:].mean() < e[:8].mean(), (e[:8], e[-8:])

"""Generate the golden fixtures in this directory from the CPU oracle.

The reference (JAX + pyscf) cannot run in this image, so these vectors pin the
*oracle* (and through it the CUDA path) against silent drift; they are not outputs of
the reference itself ("parity unpinned", see oracle/deepsolid_oracle.py).

    python tests/golden/make_golden.py
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np, torch
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O

torch.set_num_threads(os.cpu_count() or 8)
CASES = [("h4", 4, 3), ("lih_prim", 4, 3), ("graphene8", 3, 2), ("h10", 3, 2), ("li24", 2, 1)]

for name, B, steps in CASES:
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec)
    P = O.params_to_torch(pn)
    X = torch.as_tensor(C.init_walkers(sc, B, seed=4242))
    f_pd = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    f_ld = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    f_sl = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet")
    el = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    ew = O.EwaldSum(sc)
    logabs, phase, ke, ee, ei = [], [], [], [], []
    for x in X:
        s, l = f_pd(P, x)
        logabs.append(float(l)); phase.append(float(torch.angle(s)))
        k, _ = el(P, x)
        ke.append(complex(k))
        a, b, ii = ew.energy(x)
        ee.append(float(a)); ei.append(float(b))
    g = torch.Generator().manual_seed(99)
    N3 = X.shape[1]
    xi = torch.randn(steps, B, N3, generator=g, dtype=torch.float64)
    u = torch.rand(steps, B, generator=g, dtype=torch.float64)
    mc = O.make_mcmc_step(lambda p, xx: O.batch_apply(f_sl, p, xx), B, sc.lattice_vectors(), steps=steps)
    xn, pmove, masks = mc(P, X, (xi, u), 0.25)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), x=X.numpy(), logabs=np.array(logabs), phase=np.array(phase),
                        ke=np.array(ke), ee=np.array(ee), ei=np.array(ei), ii=float(ii), xi=xi.numpy(), u=u.numpy(),
                        width=0.25, x_new=xn.numpy(), masks=masks.numpy(), pmove=float(pmove), param_seed=888,
                        walker_seed=4242)
    print(name, "ke", ke[0], "ewald", ee[0] + ei[0] + float(ii), "pmove", float(pmove))

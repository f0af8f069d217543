"""Write golden vectors from the REAL bytedance/DeepSolid (JAX + pyscf) for the local-energy hot path.

This image has neither jax nor pyscf (nor a network), so this script cannot run here and no
``tests/golden/reference_*.npz`` is committed: parity stays "unpinned by reference outputs".  A maintainer
with the reference's environment closes the gap with

    pip install jax==0.2.26 jaxlib==0.1.75 pyscf chex optax ml_collections      # DeepSolid setup.py:19-27
    python tests/golden/make_reference_golden.py --reference /path/to/DeepSolid --out tests/golden

after which ``tests/test_reference_golden.py`` compares the oracle (CPU) and the CUDA path (GPU) with the file:
log|psi|, phase, kinetic energy, the three Ewald terms, the orbital matrices, and the Metropolis accept masks for
recorded noise.  Everything the consumer needs is INSIDE the file (lattice, atoms, charges, S, k-list, parameters,
walkers, noise), so it needs neither jax nor pyscf.

What is run (all reference code, unmodified; file:line relative to DeepSolid/):
  cells      test/test_cell.py:11-25 (LiH, sto-3g), supercell.get_supercell (supercell.py:64-95), S = I and diag(2,1,1)
  k-list     hf.SCF(simulation_cell, twist).init_scf().klist (hf.py:43-104), as test/test_network.py:32-34 does
  params     network.make_solid_fermi_net(...).init(key) (network.py:609-667, base_config defaults, 8 determinants)
  walkers    init_guess.init_electrons (init_guess.py:27-80), then `--burn` reference Metropolis moves
  outputs    eval_phase_and_slogdet / eval_logdet / eval_mats (network.py:563-606),
             hamiltonian.local_energy_seperate in modes for / partition / dim_batch (hamiltonian.py:194-228),
             ewaldsum.EwaldSum.energy (ewaldsum.py:185-191),
             qmc.mh_update step by step (qmc.py:153-224) with the normal / uniform draws of its own key splits recorded.

``--backend oracle`` writes the same file from the repo's CPU oracle instead: it exists ONLY so that the CPU test
suite can exercise the writer / reader plumbing (the file is stamped source="oracle-selftest" and the parity test
refuses to treat such a file as a reference pin).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = {"reference_lih_s111": np.eye(3), "reference_lih_s211": np.diag([2.0, 1.0, 1.0])}


def _flatten(params):
    """Reference pytree -> {name: array} with names 'single/0/w', ... (network.py:135-184)."""
    out = {}
    for group in ("single", "double", "orbital", "envelope"):
        for i, d in enumerate(params[group]):
            for leaf, v in d.items():
                out[f"param/{group}/{i}/{leaf}"] = np.asarray(v, dtype=np.float64)
    return out


def run_reference(S, batch, steps, burn, seed, ref_path):
    if ref_path:
        sys.path.insert(0, ref_path)
        sys.path.insert(0, os.path.join(ref_path, "test"))
    import jax
    import jax.numpy as jnp
    jax.config.update("jax_enable_x64", True)
    from pyscf.pbc import gto
    from DeepSolid import base_config, distance, ewaldsum, hamiltonian, hf, init_guess, network, qmc, supercell

    # test/test_cell.py:11-25
    cell = gto.Cell()
    L = 2 / 0.529177
    cell.atom = f"""
    Li 0 0 0
    H {L/2} {L/2} {L/2}
    """
    cell.basis = "sto-3g"
    cell.a = (1 - np.eye(3)) * L / 2
    cell.unit = "B"
    cell.verbose = 0
    cell.spin = 0
    cell.exp_to_discard = 0.1
    cell.build()
    simulation_cell = supercell.get_supercell(cell, S=S)

    twist = jnp.zeros(3)
    scf_approx = hf.SCF(simulation_cell, twist=twist)            # test/test_network.py:32-34
    scf_approx.init_scf()
    klist = [np.asarray(k, dtype=np.float64) for k in scf_approx.klist]

    cfg = base_config.default()
    cfg.network.detnet.determinants = 8
    system_dict = {"klist": scf_approx.klist, "simulation_cell": simulation_cell}
    system_dict.update(cfg.network.detnet)
    system_dict["envelope_type"] = "isotropic"
    system_dict["full_det"] = False
    nets = {m: network.make_solid_fermi_net(**system_dict, method_name=m)
            for m in ("eval_logdet", "eval_slogdet", "eval_phase_and_slogdet", "eval_mats")}

    key = jax.random.PRNGKey(seed)
    key, k_init, k_par, k_burn, k_mc = jax.random.split(key, 5)
    internal = init_guess.pyscf_to_cell(simulation_cell)
    data = init_guess.init_electrons(k_init, internal, simulation_cell.a, simulation_cell.nelec, batch_size=batch)
    params = nets["eval_logdet"].init(k_par, data=None)

    batch_slog = jax.vmap(nets["eval_slogdet"].apply, in_axes=(None, 0))
    latvec = simulation_cell.lattice_vectors()
    width = 0.15
    if burn:
        burn_step = qmc.make_mcmc_step(batch_slog, batch, latvec=latvec, steps=burn)
        data, _ = burn_step(params, data, k_burn, width)
    x0 = np.asarray(data, dtype=np.float64)

    out = {}
    sign, slog = jax.vmap(nets["eval_phase_and_slogdet"].apply, in_axes=(None, 0))(params, data)
    out["logabs"] = np.asarray(slog, dtype=np.float64)
    out["phase"] = np.asarray(jnp.angle(sign), dtype=np.float64)
    out["logdet"] = np.asarray(jax.vmap(nets["eval_logdet"].apply, in_axes=(None, 0))(params, data), dtype=np.complex128)
    mats = jax.vmap(nets["eval_mats"].apply, in_axes=(None, 0))(params, data)
    for s, m in enumerate(mats):
        out[f"mats{s}"] = np.asarray(m, dtype=np.complex128)
    for mode in ("for", "partition", "dim_batch"):
        el = hamiltonian.local_energy_seperate(nets["eval_logdet"].apply, simulation_cell, mode=mode, partition_number=3)
        ke, ew = jax.vmap(el, in_axes=(None, 0))(params, data)
        out[f"ke_{mode}"] = np.asarray(ke, dtype=np.complex128)
        out[f"ewald_{mode}"] = np.asarray(ew, dtype=np.float64)
    ewald = ewaldsum.EwaldSum(simulation_cell)
    ee, ei, ii = jax.vmap(ewald.energy)(data)
    out["ee"], out["ei"], out["ii"] = (np.asarray(v, dtype=np.float64) for v in (ee, ei, ii))
    out["ewald_alpha"] = np.float64(ewald.alpha)
    out["ewald_ng"] = np.int64(ewald.gweight.shape[0])

    # Metropolis: qmc.mh_update, one call per move, with the draws of its two key splits recorded (qmc.py:189-218)
    x1 = data
    lp = 2.0 * batch_slog(params, x1)
    xi, u, masks = [], [], []
    k = k_mc
    nacc = 0.0
    for _ in range(steps):
        k_a, sub_a = jax.random.split(k)
        xi.append(np.asarray(jax.random.normal(sub_a, shape=x1.shape), dtype=np.float64))
        _, sub_b = jax.random.split(k_a)
        u.append(np.asarray(jax.random.uniform(sub_b, shape=lp.shape), dtype=np.float64))
        x_new, k, lp_new, nacc = qmc.mh_update(params, batch_slog, x1, k, lp, nacc, latvec, stddev=width)
        masks.append(np.asarray(lp_new != lp) | np.any(np.asarray(x_new != x1), axis=-1))
        x1, lp = x_new, lp_new
    # the jitted driver must land on the same walkers (qmc.py:335-362)
    x_drv, pmove = qmc.make_mcmc_step(batch_slog, batch, latvec=latvec, steps=steps)(params, data, k_mc, width)
    assert np.allclose(np.asarray(x_drv), np.asarray(x1), atol=1e-12), "step-by-step replay differs from mcmc_step"
    out.update(xi=np.stack(xi), u=np.stack(u), masks=np.stack(masks), x_new=np.asarray(x1, dtype=np.float64),
               pmove=np.float64(pmove), width=np.float64(width))

    prim = simulation_cell.original_cell
    out.update(_flatten(params))
    out.update(x=x0, S=np.asarray(S, dtype=np.float64), nelec=np.asarray(simulation_cell.nelec, dtype=np.int64),
               prim_a=np.asarray(prim.lattice_vectors(), dtype=np.float64),
               prim_atoms=np.asarray(prim.atom_coords(), dtype=np.float64),
               prim_charges=np.asarray(prim.atom_charges(), dtype=np.float64),
               sim_a=np.asarray(simulation_cell.lattice_vectors(), dtype=np.float64),
               sim_atoms=np.asarray(simulation_cell.atom_coords(), dtype=np.float64),
               sim_charges=np.asarray(simulation_cell.atom_charges(), dtype=np.float64),
               sim_AV=np.asarray(simulation_cell.AV), sim_BV=np.asarray(simulation_cell.BV),
               prim_AV=np.asarray(prim.AV), prim_BV=np.asarray(prim.BV),
               klist0=klist[0], klist1=klist[1], energy_nuc=np.float64(simulation_cell.energy_nuc()),
               source=np.array("reference"), versions=np.array(f"jax {jax.__version__}"))
    return out


#: fixtures written with --backend shim: system (oracle/geometry.py name: primitive cell, charges and S of the reference's
#: config files), network options (network.py:609-667), Laplacian modes, walkers, Metropolis moves, burn-in moves
SHIM_CASES = {
    # the cell of the reference's own tests (test/test_cell.py), all three Laplacian modes
    "reference_shim_lih_s111": dict(system="test_cell_lih", S=np.eye(3), batch=4, steps=3, burn=10),
    "reference_shim_lih_s211": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=4, steps=3, burn=10,
                                    total_energy=True, moves=True, observables=True, pretrain=True),
    # the structural options of make_solid_fermi_net
    "reference_shim_lih_tri": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=1, burn=4,
                                   opts=dict(distance_type="tri"), modes=("for",)),
    "reference_shim_lih_diagenv": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=1, burn=4,
                                       opts=dict(envelope_type="diagonal"), modes=("for",)),
    "reference_shim_lih_fullenv": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=1, burn=4,
                                       opts=dict(envelope_type="full"), modes=("for",)),
    "reference_shim_lih_fulldet": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=1, burn=4,
                                       opts=dict(full_det=True), modes=("for",), pretrain=True),
    "reference_shim_lih_bias": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=1, burn=4,
                                    opts=dict(bias_orbitals=True), modes=("for",)),
    "reference_shim_lih_lastlayer": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=1, burn=4,
                                         opts=dict(use_last_layer=True), modes=("for",)),
    # spin-polarised cell (n_up != n_dn: network.py:322-323 splits, per-spin orbital blocks) and twisted k-points
    "reference_shim_lih_spin": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=2, burn=4,
                                    spin=2, modes=("for",)),
    "reference_shim_lih_twist": dict(system="test_cell_lih", S=np.diag([2.0, 1.0, 1.0]), batch=2, steps=2, burn=4,
                                     twist=(0.13, -0.21, 0.34), modes=("for",)),
    # minimal-image branches no config file reaches (distance.py:41-59, 91-108): an orthogonal lattice
    # that is not diagonal, and an obtuse one that the reference ALSO classifies as orthogonal (dot < tol without abs)
    "reference_shim_ortho": dict(custom=dict(a=[[3.0, 3.0, 0.0], [-2.0, 2.0, 0.0], [0.0, 0.0, 5.0]]), S=np.diag([2.0, 1.0, 1.0]),
                                 batch=3, steps=2, burn=4, modes=("for",), init_width=1.5),
    "reference_shim_obtuse": dict(custom=dict(a=[[4.0, 0.0, 0.0], [-1.9, 4.2, 0.0], [-1.5, -1.8, 5.0]]), S=np.diag([2.0, 1.0, 1.0]),
                                  batch=3, steps=2, burn=4, modes=("for",), init_width=1.5),
    # BASELINE.json configurations (SURVEY 8 d), one or two walkers each at full size
    "reference_shim_h10": dict(system="h10", batch=2, steps=2, burn=4, modes=("for", "partition")),
    "reference_shim_li24": dict(system="li24", batch=2, steps=1, burn=2, modes=("for",)),
    "reference_shim_graphite54": dict(system="graphite54", batch=2, steps=1, burn=2, modes=("for",)),
    "reference_shim_diamond64": dict(system="diamond64", batch=1, steps=1, burn=2, modes=("partition",)),
    "reference_shim_lih108": dict(system="lih108", batch=1, steps=1, burn=1, modes=("partition",)),
}
_SYMBOL_OF_Z = {1.0: "H", 2.0: "He", 3.0: "Li", 4.0: "Be", 6.0: "C"}


def run_shim(case, seed):
    """The reference's SOURCE FILES (imported unmodified from /root/reference) executed on a torch-backed stand-in for
    jax / jax.numpy / pyscf.pbc.gto (tests/golden/torch_jax_shim.py): every formula of network.py, hamiltonian.py,
    ewaldsum.py, distance.py, supercell.py and qmc.mh_update is the reference's; the array library, the autodiff
    transforms and the cell container are substitutes.  Inputs that the reference takes from pyscf / its RNG (primitive
    cell of the config file, k-point occupation of the HF solution, parameter draws, initial walkers, Metropolis noise)
    are generated here with numpy and stored in the file, so the consumer sees exactly what the reference code saw."""
    import json
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import torch
    import torch_jax_shim as shim
    shim.install()
    import jax
    from DeepSolid import ewaldsum, hamiltonian, network, qmc, supercell
    from oracle import deepsolid_oracle as O          # parameter shapes / draws only (inputs, not outputs)
    from oracle import geometry as G                  # the primitive cells of the reference's config files (inputs)

    batch, steps, burn = case["batch"], case["steps"], case["burn"]
    opts = dict(case.get("opts", {}))
    modes = tuple(case.get("modes", ("for", "partition", "dim_batch")))
    if "custom" in case:              # two He atoms at fixed fractional positions in the given lattice
        lat = np.asarray(case["custom"]["a"], dtype=np.float64)
        frac = np.array([[0.1, 0.15, 0.2], [0.6, 0.55, 0.7]])
        prim0 = G.RefCell(lat, [("He", xyz) for xyz in frac @ lat], {"He": 2.0}, name="custom")
        S = np.asarray(case["S"], dtype=np.float64)
    else:
        named = G.build_system(case["system"])
        prim0 = named.original_cell
        S = np.asarray(case.get("S", named.S), dtype=np.float64)
    # primitive cell -> the stand-in for pyscf.pbc.gto.Cell; the SUPERCELL is built by the reference (supercell.py:64-95)
    cell = shim.Cell()
    syms, charges = [], {}
    for i, ((sym, xyz), z) in enumerate(zip(prim0._atom, prim0.atom_charges())):
        syms.append(sym)
        charges[sym] = float(z)
    cell.atom = [(sym, tuple(xyz)) for sym, (_, xyz) in zip(syms, prim0._atom)]
    cell.ecp = charges if any(abs(z - shim._Z.get(sym, -1)) > 1e-12 for sym, z in charges.items()) else None
    cell.a = np.asarray(prim0.lattice_vectors(), dtype=np.float64)
    cell.unit = "B"
    cell.spin = int(case.get("spin", prim0.nelec[0] - prim0.nelec[1]))
    cell.basis, cell.exp_to_discard = "sto-3g", 0.1
    cell.build()
    simulation_cell = supercell.get_supercell(cell, S=S)

    # k-points of the supercell (supercell.py:32-48, reference code); occupation: the lowest-index k-points take the
    # remainder (the HF solution that orders them in the reference run is an input, not part of the hot path)
    kpts = np.asarray(supercell.get_supercell_kpts(simulation_cell), dtype=np.float64)
    if "twist" in case:               # twisted boundary conditions: every k-point shifted by twist . b (base_config.py:140, hf.py:60-66)
        kpts = kpts + np.asarray(case["twist"], dtype=np.float64) @ simulation_cell.reciprocal_vectors()
    klist = []
    for ns in simulation_cell.nelec:
        per, rem = divmod(ns, len(kpts))
        rows = [kpts[i] for i in range(len(kpts)) for _ in range(per + (1 if i < rem else 0))]
        klist.append(np.stack(rows) if rows else np.zeros((0, 3)))

    kw = dict(envelope_type="isotropic", bias_orbitals=False, use_last_layer=False, klist=klist,
              simulation_cell=simulation_cell, full_det=False, hidden_dims=((256, 32),) * 3, determinants=8,
              after_determinants=1, distance_type="nu")
    kw.update(opts)
    nets = {m: network.make_solid_fermi_net(**kw, method_name=m)
            for m in ("eval_logdet", "eval_slogdet", "eval_phase_and_slogdet", "eval_mats")}
    # parameters: shapes / distributions of network.py:60-186 drawn with numpy from `seed` (the file stores the seed, the
    # consumer redraws them: 0.5 M doubles would make a multi-MB fixture); everything else from a second stream
    init_kw = {k: kw[k] for k in ("envelope_type", "bias_orbitals", "use_last_layer", "full_det", "distance_type")}
    params = O.params_to_torch(O.init_params(np.random.default_rng(seed), cell.natm, simulation_cell.nelec, **init_kw))
    rng = np.random.default_rng(seed + 1)
    latvec = simulation_cell.lattice_vectors()
    # walkers: electrons on atoms + noise (init_guess.py:69-80 does the same with its own key), wrapped into the cell
    sim_atoms = simulation_cell.atom_coords()
    n_up, n_dn = simulation_cell.nelec
    idx = [i % len(sim_atoms) for i in range(n_up)] + [i % len(sim_atoms) for i in range(n_dn)]
    x = sim_atoms[idx][None] + float(case.get("init_width", 0.8)) * rng.standard_normal((batch, n_up + n_dn, 3))
    frac = x @ np.linalg.inv(latvec)
    data = torch.as_tensor(((frac - np.floor(frac)) @ latvec).reshape(batch, -1))

    def batched(f):
        return lambda p, xs: torch.stack([f(p, xs[b]) for b in range(xs.shape[0])])

    batch_slog = batched(nets["eval_slogdet"].apply)
    width = 0.15
    nacc = 0.0
    lp = 2.0 * batch_slog(params, data)
    key = jax.random.PRNGKey(seed)
    for _ in range(burn):                                     # burn-in with the reference's own update (qmc.py:153-224)
        shim.set_random_queue([rng.standard_normal(tuple(data.shape)), rng.random(tuple(lp.shape))])
        data, key, lp, nacc = qmc.mh_update(params, batch_slog, data, key, lp, nacc, latvec, stddev=width)
    x0 = data.numpy().copy()

    out = {}
    res = [nets["eval_phase_and_slogdet"].apply(params, data[b]) for b in range(batch)]
    out["logabs"] = np.asarray([float(r[1]) for r in res])
    out["phase"] = np.asarray([float(torch.angle(r[0])) for r in res])
    out["logdet"] = np.asarray([complex(nets["eval_logdet"].apply(params, data[b])) for b in range(batch)])
    mats = [nets["eval_mats"].apply(params, data[b]) for b in range(batch)]
    for s in range(len(mats[0])):
        out[f"mats{s}"] = np.stack([m[s].numpy() for m in mats]).astype(np.complex128)
    pn = int(case.get("partition_number", 3))
    for mode in modes:
        el = hamiltonian.local_energy_seperate(nets["eval_logdet"].apply, simulation_cell, mode=mode, partition_number=pn)
        kes, ews = zip(*[el(params, data[b]) for b in range(batch)])
        out[f"ke_{mode}"] = np.asarray([complex(k) for k in kes])
        out[f"ewald_{mode}"] = np.asarray([float(e) for e in ews])
    if case.get("total_energy", False):
        # train.make_loss(...).total_energy forward (train.py:37-89): loss, variance (mean of local variances quirk is a
        # multi-device matter: one device here), imaginary part and the per-walker aux
        from DeepSolid import train
        shim.LOOP_VMAP = True
        try:
            te = train.make_loss(nets["eval_logdet"].apply, None, simulation_cell, clip_local_energy=5.0,
                                 clip_type="real", mode=modes[0], partition_number=pn)
            loss, aux = te(params, data)
        finally:
            shim.LOOP_VMAP = False
        out.update(te_loss=np.float64(float(loss)), te_variance=np.float64(float(aux.variance)),
                   te_imaginary=np.float64(float(aux.imaginary)),
                   te_local_energy=aux.local_energy.numpy().astype(np.complex128))
        # the energy-gradient estimator: the reference's custom JVP rule (train.py:90-142) applied to a fixed parameter
        # tangent T gives sum_leaves <grad, T>; T = a second parameter-shaped numpy draw (seed stored)
        tangent = O.params_to_torch(O.init_params(np.random.default_rng(seed + 7), cell.natm, simulation_cell.nelec, **init_kw))
        batch_logdet = batched(nets["eval_logdet"].apply)
        for clip_type in ("real", "complex"):
            shim.LOOP_VMAP = True
            try:
                te = train.make_loss(nets["eval_logdet"].apply, batch_logdet, simulation_cell, clip_local_energy=5.0,
                                     clip_type=clip_type, mode=modes[0], partition_number=pn)
                _, (tdot, _) = te.jvp_rule((params, data), (tangent, torch.zeros_like(data)))
            finally:
                shim.LOOP_VMAP = False
            out[f"te_jvp_{clip_type}"] = np.float64(float(tdot))
        out["tangent_seed"] = np.int64(seed + 7)
    ewald = ewaldsum.EwaldSum(simulation_cell)
    parts = [ewald.energy(data[b]) for b in range(batch)]
    out["ee"], out["ei"], out["ii"] = (np.asarray([float(p[i]) for p in parts]) for i in range(3))
    out["ewald_alpha"] = np.float64(ewald.alpha)
    out["ewald_ng"] = np.int64(ewald.gweight.shape[0])

    x1 = data
    lp = 2.0 * batch_slog(params, x1)
    xi, u, masks = [], [], []
    nacc = 0.0
    for _ in range(steps):
        xi.append(rng.standard_normal(tuple(x1.shape)))
        u.append(rng.random(tuple(lp.shape)))
        shim.set_random_queue([xi[-1], u[-1]])
        x_new, key, lp_new, nacc = qmc.mh_update(params, batch_slog, x1, key, lp, nacc, latvec, stddev=width)
        masks.append((lp_new != lp).numpy() | np.any((x_new != x1).numpy(), axis=-1))
        x1, lp = x_new, lp_new
    out.update(xi=np.stack(xi), u=np.stack(u), masks=np.stack(masks), x_new=x1.numpy().astype(np.float64),
               pmove=np.float64(float(nacc) / (steps * batch)), width=np.float64(width))     # qmc.py:360

    if case.get("moves", False):
        # secondary samplers through the reference's own driver (qmc.py:290-364): one-electron Metropolis moves
        # (mh_one_electron_update, qmc.py:227-287) and drift-diffusion importance sampling (importance_update +
        # limdrift, qmc.py:63-150); the driver only returns the final walkers and pmove, which is what is stored
        nel = n_up + n_dn
        oe_steps, imp_steps = 1, 2
        oe_xi = rng.standard_normal((oe_steps * nel, batch, 1, 3))
        oe_u = rng.random((oe_steps * nel, batch))
        shim.set_random_queue([a for pair in zip(oe_xi, oe_u) for a in pair])
        step = qmc.make_mcmc_step(batch_slog, batch, latvec=latvec, steps=oe_steps, one_electron_moves=True)
        oe_x, oe_p = step(params, data, key, 0.3)
        imp_xi = rng.standard_normal((imp_steps, batch, 3 * nel))
        imp_u = rng.random((imp_steps, batch))
        shim.set_random_queue([a for pair in zip(imp_xi, imp_u) for a in pair])
        shim.LOOP_VMAP = True
        try:
            step = qmc.make_mcmc_step(batch_slog, batch, latvec=latvec, steps=imp_steps,
                                      importance_sampling=nets["eval_slogdet"].apply)
            imp_x, imp_p = step(params, data, key, 0.2)
        finally:
            shim.LOOP_VMAP = False
        out.update(oe_xi=oe_xi[:, :, 0, :], oe_u=oe_u, oe_x_new=oe_x.numpy().astype(np.float64), oe_pmove=np.float64(float(oe_p)),
                   oe_width=np.float64(0.3), imp_xi=imp_xi, imp_u=imp_u, imp_x_new=imp_x.numpy().astype(np.float64),
                   imp_pmove=np.float64(float(imp_p)), imp_width=np.float64(0.2))

    if case.get("pretrain", False):
        # pretraining loss and its parameter gradient (pretrain.py:43-107) through the reference's make_pretrain_step: the
        # optimiser handed in only records the search direction (zero update), so the step returns the loss at `params`
        from DeepSolid import pretrain as rpre
        full_det = bool(opts.get("full_det", False))
        target = [torch.as_tensor(rng.standard_normal((batch, ns, ns)) + 1j * rng.standard_normal((batch, ns, ns)))
                  for ns in simulation_cell.nelec]
        captured = {}

        class RecordingOptimizer:
            def update(self, grads, state, params=None):
                captured["g"] = grads
                return torch.utils._pytree.tree_map(torch.zeros_like, grads), state

        def batch_mats(p, xs):
            per = [nets["eval_mats"].apply(p, xs[b]) for b in range(xs.shape[0])]
            return [torch.stack([m[s] for m in per]) for s in range(len(per[0]))]

        pstep = rpre.make_pretrain_step(batch_mats, batch_slog, latvec, RecordingOptimizer(), full_det=full_det)
        shim.set_random_queue([rng.standard_normal(tuple(data.shape)), rng.random((batch,))])
        _, _, _, pt_loss, _, _ = pstep(data, target, params, None, key)
        tangent = O.params_to_torch(O.init_params(np.random.default_rng(seed + 7), cell.natm, simulation_cell.nelec, **init_kw))
        gl, tl = O._leaves(captured["g"]), O._leaves(tangent)
        out.update(pt_target0=target[0].numpy(), pt_target1=target[1].numpy(), pt_loss=np.float64(float(pt_loss)),
                   pt_dot=np.float64(sum(float((a * b).sum()) for a, b in zip(gl, tl))),
                   pt_norms=np.asarray([float(a.norm()) for a in gl]), tangent_seed=np.int64(seed + 7))

    if case.get("observables", False):              # estimator.py:15-85 on the walkers of the file
        from DeepSolid import estimator
        out["obs_sk"] = estimator.make_structure_factor(simulation_cell, nq=3)(data).numpy().astype(np.float64)
        out["obs_pol"] = np.asarray([complex(estimator.make_complex_polarization(simulation_cell, direction=d)(data))
                                     for d in range(3)])

    prim = simulation_cell.original_cell
    flat = _flatten({g: [{k: v.numpy() for k, v in d.items()} for d in params[g]] for g in params})
    out.update(param_seed=np.int64(seed), param_checksum=np.float64(sum(float(np.abs(v).sum()) for v in flat.values())))
    out.update(x=x0, S=np.asarray(S, dtype=np.float64), nelec=np.asarray(simulation_cell.nelec, dtype=np.int64),
               prim_a=np.asarray(prim.lattice_vectors(), dtype=np.float64),
               prim_atoms=np.asarray(prim.atom_coords(), dtype=np.float64),
               prim_charges=np.asarray(prim.atom_charges(), dtype=np.float64),
               sim_a=np.asarray(simulation_cell.lattice_vectors(), dtype=np.float64),
               sim_atoms=np.asarray(simulation_cell.atom_coords(), dtype=np.float64),
               sim_charges=np.asarray(simulation_cell.atom_charges(), dtype=np.float64),
               sim_AV=np.asarray(simulation_cell.AV), sim_BV=np.asarray(simulation_cell.BV),
               prim_AV=np.asarray(prim.AV), prim_BV=np.asarray(prim.BV),
               klist0=klist[0], klist1=klist[1], energy_nuc=np.float64(float(ewald.ion_ion + ewald.ii_const)),
               opts=np.array(json.dumps(opts)), modes=np.array(",".join(modes)), partition_number=np.int64(pn),
               source=np.array("reference-source/torch-shim"),
               versions=np.array(f"torch {torch.__version__} stand-in for jax; DeepSolid sources from /root/reference"))
    return out


def run_oracle(S, batch, steps, burn, seed):
    """Plumbing self-test only: the same file layout written from the CPU oracle."""
    sys.path.insert(0, ROOT)
    import torch
    from oracle import deepsolid_oracle as O, geometry as G
    L = 2 / 0.529177
    prim = G.RefCell((1 - np.eye(3)) * L / 2, [("Li", [0, 0, 0]), ("H", [L / 2] * 3)], {"Li": 3.0, "H": 1.0})
    sc = G.get_supercell(prim, S)
    klist = G.make_klist(sc)
    rng = np.random.default_rng(seed)
    pn = O.init_params(rng, prim.natm, sc.nelec)
    P = O.params_to_torch(pn)
    X = torch.as_tensor(G.init_walkers(sc, batch, seed=seed))
    nets = {m: O.make_solid_fermi_net(klist, sc, method_name=m)
            for m in ("eval_logdet", "eval_slogdet", "eval_phase_and_slogdet", "eval_mats")}
    bslog = lambda p, x: O.batch_apply(nets["eval_slogdet"], p, x)
    width = 0.15
    if burn:
        X, _, _ = O.make_mcmc_step(bslog, batch, sc.lattice_vectors(), steps=burn)(
            P, X, (torch.as_tensor(rng.standard_normal((burn,) + tuple(X.shape))), torch.as_tensor(rng.uniform(size=(burn, batch)))), width)
    out = {"logabs": [], "phase": [], "logdet": [], "mats0": [], "mats1": [], "ee": [], "ei": [], "ii": []}
    els = {m: O.local_energy_seperate(nets["eval_logdet"], sc, mode=m, partition_number=3) for m in ("for", "partition", "dim_batch")}
    for m in els:
        out[f"ke_{m}"], out[f"ewald_{m}"] = [], []
    ew = O.EwaldSum(sc)
    for b in range(batch):
        sign, slog = nets["eval_phase_and_slogdet"](P, X[b])
        out["logabs"].append(float(slog)); out["phase"].append(float(torch.angle(sign)))
        out["logdet"].append(complex(nets["eval_logdet"](P, X[b])))
        mats = nets["eval_mats"](P, X[b])
        out["mats0"].append(mats[0].numpy()); out["mats1"].append(mats[1].numpy())
        for m, el in els.items():
            ke, e = el(P, X[b])
            out[f"ke_{m}"].append(complex(ke)); out[f"ewald_{m}"].append(float(e))
        ee, ei, ii = ew.energy(X[b])
        out["ee"].append(float(ee)); out["ei"].append(float(ei)); out["ii"].append(float(ii))
    out = {k: np.asarray(v) for k, v in out.items()}
    xi = rng.standard_normal((steps,) + tuple(X.shape)); u = rng.uniform(size=(steps, batch))
    xn, pmove, masks = O.make_mcmc_step(bslog, batch, sc.lattice_vectors(), steps=steps)(P, X, (torch.as_tensor(xi), torch.as_tensor(u)), width)
    out.update(xi=xi, u=u, masks=np.asarray(masks).astype(bool), x_new=xn.numpy(), pmove=np.float64(pmove), width=np.float64(width))
    out.update(_flatten(pn))
    out.update(x=X.numpy(), S=np.asarray(S, dtype=np.float64), nelec=np.asarray(sc.nelec, dtype=np.int64),
               prim_a=prim.a, prim_atoms=prim.atom_coords(), prim_charges=prim.atom_charges(), sim_a=sc.a,
               sim_atoms=sc.atom_coords(), sim_charges=sc.atom_charges(), sim_AV=sc.AV, sim_BV=sc.BV,
               prim_AV=prim.AV, prim_BV=prim.BV, klist0=klist[0], klist1=klist[1],
               ewald_alpha=np.float64(ew.alpha), ewald_ng=np.int64(ew.gweight.shape[0]),
               energy_nuc=np.float64(ew.ion_ion + ew.ii_const),
               source=np.array("oracle-selftest"), versions=np.array("oracle"))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--reference", default=os.environ.get("DEEPSOLID_PATH", ""), help="checkout of bytedance/DeepSolid")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--backend", choices=["reference", "shim", "oracle"], default="reference")
    ap.add_argument("--batch", type=int, default=6)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--burn", type=int, default=30)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    a = ap.parse_args(argv)
    os.makedirs(a.out, exist_ok=True)
    if a.backend == "shim" and a.cases == list(CASES):
        a.cases = list(SHIM_CASES)
    for name in a.cases:
        if a.backend == "shim":
            data = run_shim(SHIM_CASES[name], a.seed)
        elif a.backend == "reference":
            data = run_reference(CASES[name], a.batch, a.steps, a.burn, a.seed, a.reference)
        else:
            data = run_oracle(CASES[name], a.batch, a.steps, a.burn, a.seed)
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"wrote {path}: {len(data)} arrays, source={data['source']}")


if __name__ == "__main__":
    main()

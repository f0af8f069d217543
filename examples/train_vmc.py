"""A VMC optimisation loop on the B200 hot path, in the shape of the reference driver's inner loop
(DeepSolid/process.py:249-372): burn-in, then per iteration one Metropolis sweep, one KFAC step on the energy
gradient, the statistics line, the move-width adaptation and (optionally) a checkpoint.  Configs, pyscf cells, HF
pretraining and the writers of the reference driver are outside the hot path; systems come from deepsolid_b200.cell.

    python examples/train_vmc.py --system h4 --batch 512 --iterations 50
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_vmc.py --system h10      # walker-sharded
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def run(system="h4", batch=512, iterations=50, burn_in=20, mcmc_steps=10, lr=5e-2, damping=1e-3, norm_constraint=1e-3,
        clip_el=5.0, move_width=0.02, adapt_frequency=10, seed=888, ckpt_dir=None, log=print, structure_factor=False):
    from deepsolid_b200 import cell as C, checkpoint, dist, estimator, hamiltonian, kfac, network, qmc, train  # noqa: F401
    import torch.distributed as td

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not td.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group("nccl")
    sc = C.build_system(system)
    kl = C.make_klist(sc)
    params = network.init_solid_fermi_net_params(seed, atoms=sc.original_cell.atom_coords(), spins=sc.nelec)
    slog = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                        determinants=8, method_name="eval_slogdet")
    logdet = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                          determinants=8, method_name="eval_logdet")
    hp = logdet.apply.hotpath()
    per_dev = batch // world
    data = torch.as_tensor(C.init_walkers(sc, per_dev, seed=666 + rank)).to(hp.tdev)
    mcmc_step = qmc.make_mcmc_step(slog.apply, per_dev, sc.a, steps=mcmc_steps)
    total_energy = train.make_loss(logdet.apply, None, sc, clip_local_energy=clip_el, clip_type="real", mode="for")
    opt = kfac.Optimizer(total_energy.value_and_grad, hp, norm_constraint=norm_constraint)
    sk_fn = estimator.make_structure_factor(sc, nq=2, hotpath=hp) if structure_factor else None
    params = {k: [{kk: torch.as_tensor(v).to(hp.tdev) for kk, v in d.items()} for d in params[k]] for k in params}

    key = 1000 * (rank + 1)
    width = move_width
    for t in range(burn_in):
        key += 1
        data, pmove = mcmc_step(params, data, key, width)
    history, pmoves = [], np.zeros(adapt_frequency)
    t0 = time.time()
    for t in range(iterations):
        key += 1
        data, pmove = mcmc_step(params, data, key, width)
        params, stats = opt.step(params, data, learning_rate=lr, damping=damping)
        aux = stats["aux"]
        row = {"step": t, "energy": float(stats["loss"]) / sc.scale, "variance": float(aux.variance) / sc.scale ** 2,
               "pmove": float(pmove), "imaginary": float(aux.imaginary) / sc.scale,
               "kinetic": float(torch.as_tensor(aux.kinetic).real.mean()) / sc.scale,
               "ewald": float(torch.as_tensor(aux.ewald).mean()) / sc.scale}
        if sk_fn is not None:
            row["structure_factor"] = sk_fn(data).cpu().numpy()
        history.append(row)
        if rank == 0 and log is not None:
            log("Step %05d: %03.4f E_h, variance=%03.4f E_h^2, pmove=%0.2f, imaginary part=%03.4f, kinetic=%03.4f E_h, "
                "ewald=%03.4f E_h" % (t, row["energy"], row["variance"], row["pmove"], row["imaginary"], row["kinetic"],
                                      row["ewald"]))
        if t > 0 and t % adapt_frequency == 0:          # process.py:366-371
            if pmoves.mean() > 0.55:
                width *= 1.1
            if pmoves.mean() < 0.5:
                width /= 1.1
            pmoves[:] = 0
        pmoves[t % adapt_frequency] = float(pmove)
    if ckpt_dir and rank == 0:
        checkpoint.save(ckpt_dir, iterations, data.cpu().numpy()[None],
                        {k: [{kk: v.cpu().numpy() for kk, v in d.items()} for d in params[k]] for k in params},
                        mcmc_width=width)
    return history, time.time() - t0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="h4")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--burn-in", type=int, default=20)
    ap.add_argument("--lr", type=float, default=5e-2)
    ap.add_argument("--ckpt-dir", default=None)
    a = ap.parse_args()
    hist, secs = run(a.system, a.batch, a.iterations, a.burn_in, lr=a.lr, ckpt_dir=a.ckpt_dir)
    e0 = np.mean([h["energy"] for h in hist[:5]])
    e1 = np.mean([h["energy"] for h in hist[-5:]])
    if int(os.environ.get("RANK", "0")) == 0:
        print("energy per primitive cell: first 5 steps %.4f, last 5 steps %.4f E_h; %.2f s per iteration"
              % (e0, e1, secs / max(len(hist), 1)))
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        td.destroy_process_group()

"""Drop-in for ``DeepSolid.qmc.make_mcmc_step`` (qmc.py:290-364): Metropolis all-electron moves, one-electron
moves and drift-diffusion importance sampling."""
from __future__ import annotations

import torch

from . import dist as _dist


def limdrift(g: torch.Tensor, cutoff: float = 1.0) -> torch.Tensor:
    """qmc.py:63-82: cap the magnitude of every electron's drift 3-vector at ``cutoff``."""
    shape = g.shape
    g3 = g.reshape(-1, 3)
    tot = torch.linalg.norm(g3, dim=-1)
    normalize = torch.clamp(tot, min=cutoff)
    return (cutoff * g3 / normalize[:, None]).reshape(shape)


def enforce_pbc(latvec: torch.Tensor, epos: torch.Tensor) -> torch.Tensor:
    """distance.enforce_pbc (distance.py:144-163) on device tensors (importance-sampled proposals)."""
    lat = torch.as_tensor(latvec, dtype=torch.float64, device=epos.device)
    frac = epos.reshape(-1, 3) @ torch.linalg.inv(lat)
    frac = frac - torch.floor(frac)
    return (frac @ lat).reshape(epos.shape)


def make_mcmc_step(batch_slog_network, batch_per_device, latvec, steps=10, atoms=None,
                   importance_sampling=None, one_electron_moves=False):
    """Returns ``mcmc_step(params, data, key, width) -> (data, pmove)``.

    ``key`` is an int seed for the device Philox stream, or a tuple ``(xi, u)`` of
    caller-supplied noise (gaussians ``(steps,B,3N)``, uniforms ``(steps,B)``; for one-electron moves
    ``(steps*N,B,3)`` and ``(steps*N,B)``) -- the form parity tests use, since jax.random cannot be
    reproduced without JAX.  ``pmove`` is averaged over ranks when torch.distributed is initialised
    (``pmean``, qmc.py:360-361).  ``importance_sampling``: any truthy value (the reference passes the network
    to differentiate, process.py:183-189) selects the drifted proposal of ``importance_update`` (qmc.py:83-150).
    """
    if importance_sampling is not None and one_electron_moves:
        raise ValueError("Importance sampling for one elec move is not implemented yet")
    if atoms is not None:
        raise ValueError("asymmetric (harmonic-mean) proposals are not implemented in the CUDA hot path")
    getter = getattr(batch_slog_network, "hotpath", None)
    if getter is None:
        raise TypeError("batch_slog_network must be the .apply of a deepsolid_b200 network (eval_slogdet)")
    lat = torch.as_tensor(latvec, dtype=torch.float64)

    def importance_step(params, data, key, width, return_masks=False):
        """importance_update (qmc.py:83-150, symmetric branch): x2 = x1 + g + s^2 drift(x1), Green's function ratio."""
        hp = getter()
        hp.set_params(params)
        dev = hp.tdev
        x1 = torch.as_tensor(data, dtype=torch.float64).to(dev)
        B = x1.shape[0]
        if isinstance(key, (tuple, list)):
            xi, u = (torch.as_tensor(k, dtype=torch.float64).to(dev) for k in key)
        else:
            gen = torch.Generator(device=dev).manual_seed(int(key))
            xi = torch.randn(steps, B, x1.shape[1], dtype=torch.float64, device=dev, generator=gen)
            u = torch.rand(steps, B, dtype=torch.float64, device=dev, generator=gen)
        width = float(width)
        la, _, grad = hp.logpsi_grad_x(x1)
        lp_1 = 2.0 * la
        grad = limdrift(grad)
        nacc = torch.zeros((), dtype=torch.float64, device=dev)
        masks = []
        for s in range(steps):
            gauss = width * xi[s]
            x2 = enforce_pbc(lat, x1 + gauss + width ** 2 * grad)
            la2, _, new_grad = hp.logpsi_grad_x(x2)
            new_grad = limdrift(new_grad)
            forward = (gauss ** 2).sum(-1)
            backward = ((gauss + width ** 2 * (grad + new_grad)) ** 2).sum(-1)
            lp_2 = 2.0 * la2 + (forward - backward) / (2.0 * width ** 2)
            cond = (lp_2 - lp_1) > torch.log(u[s])
            x1 = torch.where(cond[:, None], x2, x1)
            lp_1 = torch.where(cond, lp_2, lp_1)
            grad = torch.where(cond[:, None], new_grad, grad)
            nacc = nacc + cond.sum()
            masks.append(cond)
        pmove = _dist.pmean(nacc / (steps * batch_per_device))
        if not torch.as_tensor(data).is_cuda:
            x1 = x1.cpu()
        if return_masks:
            return x1, pmove, torch.stack(masks).to(torch.uint8)
        return x1, pmove

    def mcmc_step(params, data, key, width, return_masks=False):
        hp = getter()
        hp.set_params(params)
        xi = u = None
        seed = 0
        if isinstance(key, (tuple, list)):
            xi, u = key
        else:
            seed = int(key)
        new, nacc, masks = hp.mcmc(data, steps, float(width), seed=seed, xi=xi, u=u, return_masks=return_masks,
                                   one_electron=bool(one_electron_moves))
        nsteps = steps * hp.nelec if one_electron_moves else steps
        pmove = nacc / (nsteps * batch_per_device)
        pmove = _dist.pmean(pmove)
        if return_masks:
            return new, pmove[0], masks
        return new, pmove[0]

    return importance_step if importance_sampling is not None else mcmc_step

"""Drop-in for ``DeepSolid.qmc.make_mcmc_step`` (qmc.py:290-364), Metropolis all-electron moves."""
from __future__ import annotations

import torch

from . import dist as _dist


def make_mcmc_step(batch_slog_network, batch_per_device, latvec, steps=10, atoms=None,
                   importance_sampling=None, one_electron_moves=False):
    """Returns ``mcmc_step(params, data, key, width) -> (data, pmove)``.

    ``key`` is an int seed for the device Philox stream, or a tuple ``(xi, u)`` of
    caller-supplied noise (gaussians ``(steps,B,3N)``, uniforms ``(steps,B)``) -- the
    form parity tests use, since jax.random cannot be reproduced without JAX.
    ``pmove`` is averaged over ranks when torch.distributed is initialised
    (``pmean``, qmc.py:360-361).
    """
    if importance_sampling is not None:
        if one_electron_moves:
            raise ValueError("Importance sampling for one elec move is not implemented yet")
        raise ValueError("importance sampling is not implemented in the CUDA hot path")
    if one_electron_moves:
        raise ValueError("one-electron moves are not implemented in the CUDA hot path")
    if atoms is not None:
        raise ValueError("asymmetric (harmonic-mean) proposals are not implemented in the CUDA hot path")
    getter = getattr(batch_slog_network, "hotpath", None)
    if getter is None:
        raise TypeError("batch_slog_network must be the .apply of a deepsolid_b200 network (eval_slogdet)")
    del latvec      # the context already holds the simulation-cell lattice (process.py:185)

    def mcmc_step(params, data, key, width, return_masks=False):
        hp = getter()
        hp.set_params(params)
        xi = u = None
        seed = 0
        if isinstance(key, (tuple, list)):
            xi, u = key
        else:
            seed = int(key)
        new, nacc, masks = hp.mcmc(data, steps, float(width), seed=seed, xi=xi, u=u, return_masks=return_masks)
        pmove = nacc / (steps * batch_per_device)
        pmove = _dist.pmean(pmove)
        if return_masks:
            return new, pmove[0], masks
        return new, pmove[0]

    return mcmc_step

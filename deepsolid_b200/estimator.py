"""Observables of DeepSolid/estimator.py on the device: the plane-wave sums come from the CUDA library (ds_rho_q), the
batch means and the cross-device mean (constants.pmean_if_pmap) are the few scalars left to the host."""
from __future__ import annotations

import numpy as np
import torch

from . import dist


def make_complex_polarization(simulation_cell, direction: int = 0, ndim: int = 3, hotpath=None):
    """estimator.py:15-40: ``<exp(i G_dir . sum_i x_i)>`` over the walkers, the order parameter of the hydrogen chain."""
    if ndim != 3:
        raise ValueError("only ndim=3 walkers are supported")
    rec_vec = np.asarray(simulation_cell.reciprocal_vectors())[direction]

    def complex_polarization(data, hp=hotpath):
        if hp is None:
            raise ValueError("complex_polarization needs the HotPath of the network (hotpath=...)")
        pol = hp.rho_q(data, rec_vec[None, :], mode=1)[:, 0]
        return _pmean_c(pol.mean(dim=0))

    return complex_polarization


def _pmean_c(z: torch.Tensor) -> torch.Tensor:
    return torch.view_as_complex(dist.pmean(torch.view_as_real(z.contiguous()).contiguous()))


def make_structure_factor(simulation_cell, nq: int = 4, ndim: int = 3, hotpath=None):
    """estimator.py:42-85: S(q) = (<|rho_q|^2> - |<rho_q>|^2) / N_e on the nq^3 mesh of supercell reciprocal vectors
    (PRB 94, 035126), q ordered as jnp.meshgrid(...)(default 'xy' indexing).ravel()."""
    if ndim != 3:
        raise ValueError("only ndim=3 walkers are supported")
    mesh = np.meshgrid(*[np.arange(nq) for _ in range(3)])
    points = np.stack([m.ravel() for m in mesh], axis=0).T
    qvecs = points @ np.asarray(simulation_cell.reciprocal_vectors())
    nelec = int(sum(simulation_cell.nelec))

    def structure_factor(data, hp=hotpath):
        if hp is None:
            raise ValueError("structure_factor needs the HotPath of the network (hotpath=...)")
        rho = hp.rho_q(data, qvecs, mode=0)
        one = _pmean_c(rho.mean(dim=0))
        two = dist.pmean((rho.abs() ** 2).mean(dim=0))
        return (two - one.abs() ** 2) / nelec

    return structure_factor

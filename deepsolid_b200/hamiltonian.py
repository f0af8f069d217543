"""Drop-in for ``DeepSolid.hamiltonian.local_energy_seperate`` (hamiltonian.py:194-228).

``f`` must be the ``apply`` of a network built by ``deepsolid_b200.network`` with
``method_name='eval_logdet'`` (as in process.py:113,191): the kinetic energy is not
obtained by differentiating ``f`` but by the fused forward-Laplacian CUDA sweep that
belongs to the same network context.
"""
from __future__ import annotations

from .hotpath import LAP_MODES


def _hotpath_of(f):
    getter = getattr(f, "hotpath", None)
    if getter is None:
        raise TypeError("f must be the .apply of a deepsolid_b200.network.make_solid_fermi_net network")
    return getter()


def local_ewald_energy(simulation_cell, f=None):
    """hamiltonian.py:163-179 -> ew(x) = ee + ei + ii."""
    if f is None:
        raise TypeError("pass the network apply whose context evaluates the Ewald sum")

    def _local_ewald_energy(x):
        ee, ei, ii = _hotpath_of(f).ewald(x)
        return ee + ei + ii
    return _local_ewald_energy


def local_energy_seperate(f, simulation_cell, mode="for", partition_number=3):
    """-> _local_energy(params, x) -> (kinetic: complex, ewald: real); natively batched."""
    if mode not in LAP_MODES:
        raise ValueError("Unrecognized laplacian evaluation mode.")
    if getattr(f, "simulation_cell", simulation_cell) is not simulation_cell:
        raise ValueError("f was built for a different simulation cell")

    def _local_energy(params, x):
        hp = _hotpath_of(f)
        hp.set_params(params)
        return hp.local_energy(x, mode=mode, partition_number=partition_number)

    _local_energy.hotpath = getattr(f, "hotpath", None)
    return _local_energy

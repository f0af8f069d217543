"""Forward half of ``DeepSolid.train.make_loss`` (train.py:37-89): total energy and its
statistics, reduced across ranks with ONE all-reduce of a packed vector."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import dist as _dist
from . import hamiltonian


@dataclass
class AuxiliaryLossData:           # train.py:28-34
    variance: torch.Tensor
    local_energy: torch.Tensor
    imaginary: torch.Tensor
    kinetic: torch.Tensor
    ewald: torch.Tensor


def reduce_energy_stats(stats6: torch.Tensor):
    """stats6 = local [sum Re e, sum Im e, sum |e|^2, sum Re ke, sum ew, n] -> (loss, imaginary, variance).

    Mirrors train.py:76-80 including its quirk: every device subtracts its *local*
    |Re mean|^2 before the pmean, i.e. variance = mean over devices of local variances.
    One all-reduce of [mean_re, mean_im, local_var, 1]."""
    n = stats6[5]
    mean_re, mean_im = stats6[0] / n, stats6[1] / n
    local_var = stats6[2] / n - mean_re.abs() ** 2
    packed = torch.stack([mean_re, mean_im, local_var, torch.ones_like(n)])
    packed = _dist.psum(packed)
    return packed[0] / packed[3], packed[1] / packed[3], packed[2] / packed[3]


def make_loss(network, batch_network, simulation_cell, clip_local_energy=5.0, clip_type="real", mode="for",
              partition_number=3):
    """Returns ``total_energy(params, data) -> (loss, AuxiliaryLossData)`` (train.py:66-89).
    The custom JVP (train.py:91-142) that turns it into a gradient estimator is not part
    of this path yet."""
    del clip_local_energy, clip_type, batch_network
    el_fun = hamiltonian.local_energy_seperate(network, simulation_cell=simulation_cell, mode=mode,
                                               partition_number=partition_number)

    def total_energy(params, data):
        ke, ew = el_fun(params, data)
        e_l = ke + ew
        hp = el_fun.hotpath()
        if ke.is_cuda:
            stats = hp.energy_stats(ke, ew)
        else:
            stats = torch.stack([e_l.real.sum(), e_l.imag.sum(), (e_l.abs() ** 2).sum(), ke.real.sum(), ew.sum(),
                                 torch.tensor(float(e_l.numel()), dtype=torch.float64)])
        loss, imaginary, variance = reduce_energy_stats(stats)
        return loss, AuxiliaryLossData(variance=variance, local_energy=e_l, imaginary=imaginary, kinetic=ke, ewald=ew)

    return total_energy

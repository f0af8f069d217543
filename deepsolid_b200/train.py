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


def clip_difference(diff: torch.Tensor, clip_local_energy: float, clip_type: str) -> torch.Tensor:
    """clip_diff of total_energy_jvp (train.py:101-127): `real` clips Re and Im separately around 0 by
    clip * mean|.|, `complex` clips the radius around its median by clip * std; statistics are pmean'ed."""
    if clip_local_energy <= 0.0:
        return diff
    if clip_type == "complex":
        radius, phase = diff.abs(), torch.angle(diff)
        radius_tv = _dist.pmean(radius.std(unbiased=False))
        radius_mean = _dist.pmean(torch.quantile(radius, 0.5, interpolation="midpoint"))     # jnp.median
        clip_radius = torch.maximum(torch.minimum(radius, radius_mean + radius_tv * clip_local_energy),
                                    radius_mean - radius_tv * clip_local_energy)
        return torch.polar(clip_radius, phase)
    if clip_type == "real":
        tv_re = _dist.pmean(diff.real.abs().mean())
        tv_im = _dist.pmean(diff.imag.abs().mean())
        re = torch.maximum(torch.minimum(diff.real, clip_local_energy * tv_re), -clip_local_energy * tv_re)
        im = torch.maximum(torch.minimum(diff.imag, clip_local_energy * tv_im), -clip_local_energy * tv_im)
        return torch.complex(re, im)
    raise ValueError("Unrecognized clip type.")


def make_loss(network, batch_network, simulation_cell, clip_local_energy=5.0, clip_type="real", mode="for",
              partition_number=3):
    """Returns ``total_energy(params, data) -> (loss, AuxiliaryLossData)`` (train.py:66-89).

    The reference turns it into a gradient estimator with a custom JVP (train.py:91-142); a torch caller has no
    tracer to hand a JVP rule to, so the same estimator is exposed as
    ``total_energy.value_and_grad(params, data) -> ((loss, aux), grads)``:
    ``grads = d/dparams mean(Re(clip_diff * conj(log psi)))`` with ``clip_diff`` held constant, computed by one
    reverse sweep of the CUDA network (``ds_logpsi_vjp``); the cross-device pmean of the gradient is left to the
    optimizer, as in the reference (kfac_ferminet_alpha optimizer.py:423)."""
    del batch_network
    if clip_type not in ("real", "complex"):
        raise ValueError("Unrecognized clip type.")
    el_fun = hamiltonian.local_energy_seperate(network, simulation_cell=simulation_cell, mode=mode,
                                               partition_number=partition_number)

    def total_energy(params, data):
        ke, ew = el_fun(params, data)
        e_l = ke + ew
        hp = el_fun.hotpath()
        if ke.is_cuda:
            stats = hp.energy_stats(ke, ew)
            comm = _dist.native_comm(hp.device)
            if comm is not None:          # N > 1 over NCCL: the reduction is the library's (ds_stats_allreduce)
                out8 = hp.stats_allreduce(stats, comm)
                return out8[0], AuxiliaryLossData(variance=out8[2], local_energy=e_l, imaginary=out8[1], kinetic=ke,
                                                  ewald=ew)
        else:
            stats = torch.stack([e_l.real.sum(), e_l.imag.sum(), (e_l.abs() ** 2).sum(), ke.real.sum(), ew.sum(),
                                 torch.tensor(float(e_l.numel()), dtype=torch.float64)])
        loss, imaginary, variance = reduce_energy_stats(stats)
        return loss, AuxiliaryLossData(variance=variance, local_energy=e_l, imaginary=imaginary, kinetic=ke, ewald=ew)

    def value_and_grad(params, data):
        loss, aux = total_energy(params, data)
        diff = aux.local_energy - loss
        clip_diff = clip_difference(diff, clip_local_energy, clip_type)
        n = clip_diff.numel()
        hp = el_fun.hotpath()
        hp.set_params(params)
        grads = hp.logpsi_vjp(data, clip_diff.real / n, clip_diff.imag / n)
        return (loss, aux), grads

    total_energy.value_and_grad = value_and_grad
    return total_energy


def make_training_step(mcmc_step, val_and_grad, opt_update):
    """train.py:147-185: one iteration = Metropolis sweep, energy and gradient, cross-rank mean of the gradient,
    optimiser update.  ``val_and_grad(params, data) -> ((loss, aux), grads)`` is ``make_loss(...).value_and_grad``;
    ``opt_update(t, grads, params, state) -> (state, params)``.  Returns
    ``step(t, data, params, state, key, mcmc_width) -> (data, params, state, loss, aux, pmove, search_direction)``."""
    from .hotpath import flatten_params, unflatten_params

    def step(t, data, params, state, key, mcmc_width):
        data, pmove = mcmc_step(params, data, key, mcmc_width)
        (loss, aux_data), search_direction = val_and_grad(params, data)
        leaves = [_dist.pmean(torch.as_tensor(g)) for g in flatten_params(search_direction)]
        search_direction = unflatten_params(leaves, len(params["single"]), "b" in params["orbital"][0], len(params["double"]) == len(params["single"]))
        state, params = opt_update(t, search_direction, params, state)
        return data, params, state, loss, aux_data, pmove, search_direction

    return step


def learning_rate_schedule(rate: float = 5e-2, decay: float = 1.0, delay: float = 10000.0):
    """process.py:200-202: ``rate * (1 / (1 + t / delay)) ** decay`` (defaults of base_config.py:46-50)."""
    return lambda t: rate * (1.0 / (1.0 + (t / delay))) ** decay


def make_adam_update(schedule=None, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """The 'adam' branch of process.py:205-208, 235-246: scale_by_adam, then the learning-rate schedule, then -1.
    Returns ``(init(params) -> state, opt_update(t, grads, params, state) -> (state, params))``."""
    from .hotpath import flatten_params, unflatten_params
    from .pretrain import Adam
    schedule = schedule or learning_rate_schedule()
    adam = Adam(1.0, b1=b1, b2=b2, eps=eps)             # unit step: the schedule supplies the rate

    def opt_update(t, grads, params, state):
        updates, state = adam.update(grads, state, params)
        lr = float(schedule(state["count"] - 1))
        leaves = [torch.as_tensor(p).to(u.device) + lr * u for p, u in zip(flatten_params(params), updates)]
        return state, unflatten_params(leaves, len(params["single"]), "b" in params["orbital"][0], len(params["double"]) == len(params["single"]))

    return adam.init, opt_update

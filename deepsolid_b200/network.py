"""Drop-in for ``DeepSolid.network.make_solid_fermi_net`` (network.py:609-667).

Same factory signature, same ``method.init`` / ``method.apply`` pair and the same
``method_name`` switch; ``apply`` runs the CUDA hot path.  The reference's callers
wrap ``apply`` in ``jax.vmap`` (process.py:116-118); here ``apply`` is natively
batched: ``x`` may be one walker ``(3N,)`` or a batch ``(B, 3N)``.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from .hotpath import HotPath

_METHODS = ("eval_slogdet", "eval_logdet", "eval_mats", "eval_phase_and_slogdet")


def init_solid_fermi_net_params(key, data=None, *, atoms, spins, envelope_type="isotropic", bias_orbitals=False,
                                use_last_layer=False, eps=0.01, full_det=False,
                                hidden_dims=((256, 32),) * 3, determinants=8, after_determinants=1,
                                distance_type="nu"):
    """Parameter pytree with the shapes and distributions of network.py:60-186.

    ``key`` seeds a numpy Generator (jax.random is not reproducible without JAX).
    Leaves are float64 torch tensors on the CPU; move them where you like.
    """
    del after_determinants, data, eps
    rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(int(key))
    natom = np.asarray(atoms).shape[0]
    if distance_type == "nu":
        in_dims = (natom * 4, 4)
    elif distance_type == "tri":
        in_dims = (natom * 7, 7)
    else:
        raise ValueError("Unrecognized distance function.")
    active = [s for s in spins if s > 0]
    nch = len(active)
    dims_one_in = ([(nch + 1) * in_dims[0] + nch * in_dims[1]] +
                   [(nch + 1) * h[0] + nch * h[1] for h in hidden_dims])
    if not use_last_layer:
        dims_one_in[-1] = hidden_dims[-1][0]
    dims_one_out = [h[0] for h in hidden_dims]
    dims_two = [in_dims[1]] + [h[1] for h in hidden_dims]
    len_double = len(hidden_dims) if use_last_layer else len(hidden_dims) - 1
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    params = {"single": [], "double": [], "orbital": [], "envelope": []}
    for s in active:
        npar = sum(spins) * determinants if full_det else s * determinants
        env = {"pi": t(np.ones((natom, npar)))}
        if envelope_type == "isotropic":
            env["sigma"] = t(np.ones((natom, npar)))
        elif envelope_type == "diagonal":
            env["sigma"] = t(np.ones((natom, 3, npar)))
        elif envelope_type == "full":
            env["sigma"] = t(np.tile(np.eye(3)[..., None, None], [1, 1, natom, npar]))
        params["envelope"].append(env)
    for i in range(len(hidden_dims)):
        params["single"].append({
            "w": t(rng.standard_normal((dims_one_in[i], dims_one_out[i])) / np.sqrt(float(dims_one_in[i]))),
            "b": t(rng.standard_normal((dims_one_out[i],)))})
        if i < len_double:
            params["double"].append({
                "w": t(rng.standard_normal((dims_two[i], dims_two[i + 1])) / np.sqrt(float(dims_two[i]))),
                "b": t(rng.standard_normal((dims_two[i + 1],)))})
    for s in active:
        npar = sum(spins) * determinants if full_det else s * determinants
        orb = {"w": t(rng.standard_normal((dims_one_in[-1], 2 * npar)) / np.sqrt(float(dims_one_in[-1])))}
        if bias_orbitals:
            orb["b"] = t(rng.standard_normal((2 * npar,)))
        params["orbital"].append(orb)
    return params


def make_solid_fermi_net(envelope_type: str = "full", bias_orbitals: bool = False, use_last_layer: bool = False,
                         klist=None, simulation_cell=None, full_det: bool = True,
                         hidden_dims=((256, 32), (256, 32), (256, 32)), determinants: int = 16,
                         after_determinants=1, distance_type="nu", method_name="eval_logdet",
                         device: Optional[int] = None, hotpath: Optional[HotPath] = None):
    """network.py:609-667.  The signature defaults are the reference's; the CUDA path
    implements the configuration the reference actually runs (base_config.py:129-139):
    envelope_type='isotropic', full_det=False, use_last_layer=False, bias_orbitals=False, with either
    distance_type ('nu' or 'tri').  Anything else raises ValueError at construction."""
    if method_name not in _METHODS:
        raise ValueError("Method name is not in class dir.")
    if distance_type not in ("nu", "tri"):
        raise ValueError("Unrecognized distance function.")
    unsupported = []
    if envelope_type not in ("isotropic", "diagonal", "full"):
        unsupported.append(f"envelope_type={envelope_type!r}")
    if use_last_layer:
        unsupported.append("use_last_layer=True")
    if unsupported:
        raise ValueError("not implemented in the CUDA hot path: " + ", ".join(unsupported) +
                         " (use_last_layer=True is the one structural option without a CUDA path)")
    if simulation_cell is None or klist is None:
        raise ValueError("simulation_cell and klist are required")
    hd = tuple(tuple(int(v) for v in h) for h in hidden_dims)
    if len({h[0] for h in hd}) != 1 or len({h[1] for h in hd}) != 1:
        raise ValueError("the CUDA hot path needs equal widths in every layer of hidden_dims")
    if not 2 <= len(hd) <= 4:
        raise ValueError("the CUDA hot path supports 2 to 4 layers")
    if hd[0][0] % 2 or hd[0][1] % 2 or not 2 <= hd[0][1] <= 32 or hd[0][0] < 2:
        raise ValueError("the CUDA hot path needs even stream widths and a two-electron width of at most 32")

    state = {"hp": hotpath}

    def _hp() -> HotPath:
        if state["hp"] is None:
            state["hp"] = HotPath(simulation_cell, klist, hidden_dims=hidden_dims, determinants=determinants,
                                  device=device, distance_type=distance_type, envelope_type=envelope_type,
                                  bias_orbitals=bias_orbitals, full_det=full_det)
        return state["hp"]

    def init(key, data=None):
        return init_solid_fermi_net_params(
            key, data, atoms=simulation_cell.original_cell.atom_coords(), spins=simulation_cell.nelec,
            envelope_type=envelope_type, bias_orbitals=bias_orbitals, use_last_layer=use_last_layer,
            full_det=full_det, hidden_dims=hidden_dims, determinants=determinants,
            after_determinants=after_determinants, distance_type=distance_type)

    def apply(params, x):
        hp = _hp()
        hp.set_params(params)
        if method_name == "eval_mats":
            return hp.orbitals(x)
        la, ph = hp.logpsi(x)
        if method_name == "eval_slogdet":
            return la
        if method_name == "eval_logdet":       # log(sign) + slog = slog + i*angle
            return torch.complex(la, ph)
        return torch.polar(torch.ones_like(ph), ph), la     # eval_phase_and_slogdet

    method = namedtuple("method", ["init", "apply"])
    apply.hotpath = _hp
    apply.method_name = method_name
    apply.simulation_cell = simulation_cell
    m = method(init=init, apply=apply)
    return m

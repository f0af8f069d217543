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


def parameter_schema(natom: int, spins, envelope_type="isotropic", bias_orbitals=False, use_last_layer=False,
                     full_det=False, hidden_dims=((256, 32),) * 3, determinants=8, distance_type="nu"):
    """The leaves of the reference parameter pytree (network.py:135-184) as a flat table, in the order their random
    numbers are drawn: ``(group, index, leaf, shape, init)`` with init ``'ones'``, ``'eye'`` (3x3 identity per atom
    and orbital), ``'bias'`` (standard normal) or ``('weight', fan_in)`` (standard normal / sqrt(fan_in))."""
    feat = {"nu": 4, "tri": 7}.get(distance_type)
    if feat is None:
        raise ValueError("Unrecognized distance function.")
    channels = sum(1 for s in spins if s > 0)
    widths = [(natom * feat, feat)] + [tuple(h) for h in hidden_dims]            # (one-electron, pair) width per level
    n_layers = len(hidden_dims)
    sigma_shape = {"isotropic": lambda q: (natom, q), "diagonal": lambda q: (natom, 3, q),
                   "full": lambda q: (3, 3, natom, q)}
    rows = []
    per_spin = [(sum(spins) if full_det else s) * determinants for s in spins if s > 0]
    for i, q in enumerate(per_spin):
        rows.append(("envelope", i, "pi", (natom, q), "ones"))
        if envelope_type in sigma_shape:
            rows.append(("envelope", i, "sigma", sigma_shape[envelope_type](q), "eye" if envelope_type == "full" else "ones"))
    for i in range(n_layers):
        one, two = widths[i]
        fan_in = (channels + 1) * one + channels * two                            # own + spin means + pair means
        rows.append(("single", i, "w", (fan_in, widths[i + 1][0]), ("weight", fan_in)))
        rows.append(("single", i, "b", (widths[i + 1][0],), "bias"))
        if i < n_layers - 1 or use_last_layer:
            rows.append(("double", i, "w", (two, widths[i + 1][1]), ("weight", two)))
            rows.append(("double", i, "b", (widths[i + 1][1],), "bias"))
    one, two = widths[-1]
    orb_in = (channels + 1) * one + channels * two if use_last_layer else one
    for i, q in enumerate(per_spin):
        rows.append(("orbital", i, "w", (orb_in, 2 * q), ("weight", orb_in)))
        if bias_orbitals:
            rows.append(("orbital", i, "b", (2 * q,), "bias"))
    return rows


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
    params = {"single": [], "double": [], "orbital": [], "envelope": []}
    for group, index, leaf, shape, init in parameter_schema(natom, spins, envelope_type, bias_orbitals, use_last_layer,
                                                            full_det, hidden_dims, determinants, distance_type):
        if init == "ones":
            value = np.ones(shape)
        elif init == "eye":
            value = np.broadcast_to(np.eye(3)[:, :, None, None], shape).copy()
        elif init == "bias":
            value = rng.standard_normal(shape)
        else:
            value = rng.standard_normal(shape) / np.sqrt(float(init[1]))
        while len(params[group]) <= index:
            params[group].append({})
        params[group][index][leaf] = torch.as_tensor(np.asarray(value, dtype=np.float64))
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
    if unsupported:
        raise ValueError("not implemented in the CUDA hot path: " + ", ".join(unsupported))
    if use_last_layer and len(hidden_dims) > 3:
        raise ValueError("use_last_layer=True is implemented for at most 3 layers (forward paths only)")
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
                                  bias_orbitals=bias_orbitals, full_det=full_det, use_last_layer=use_last_layer)
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

"""Hartree-Fock pretraining on the B200 hot path: the mirror of DeepSolid/pretrain.py.

``make_pretrain_step`` has the reference's signature and semantics (pretrain.py:43-111): the loss is the mean squared
modulus of (target orbital matrix - network orbital matrix) per spin block, averaged over blocks; its parameter gradient
is one pullback through ``eval_mats`` (``ds_orbitals_vjp``), the update is Adam (the reference uses ``optax.adam``,
pretrain.py:140), followed by one Metropolis move on the network density (pretrain.py:100-108).

The Hartree-Fock target comes from pyscf in the reference (hf.SCF.eval_orb_mat, hf.py:136-153); pyscf is not available
here, so ``pretrain_hartree_fock`` takes any object with ``eval_orb_mat(coord[B, n_e, 3]) -> [up (B,n,n), down (B,n,n)]``;
``PlaneWaveSCF`` is the free-electron determinant on the occupied k-points, used by the tests and examples."""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np
import torch

from . import dist, qmc
from .hotpath import flatten_params, unflatten_params


class Adam:
    """optax.adam(learning_rate) (b1 = 0.9, b2 = 0.999, eps = 1e-8, eps_root = 0): init / update like optax."""

    def __init__(self, learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), b1, b2, eps

    def init(self, params):
        leaves = flatten_params(params)
        z = [torch.zeros_like(torch.as_tensor(t)) for t in leaves]
        return {"count": 0, "mu": z, "nu": [t.clone() for t in z]}

    def update(self, grads, state, params=None):
        g = [torch.as_tensor(t) for t in flatten_params(grads)]
        count = state["count"] + 1
        mu = [self.b1 * m.to(t.device) + (1 - self.b1) * t for m, t in zip(state["mu"], g)]
        nu = [self.b2 * v.to(t.device) + (1 - self.b2) * t * t for v, t in zip(state["nu"], g)]
        c1, c2 = 1 - self.b1 ** count, 1 - self.b2 ** count
        upd = [-self.lr * (m / c1) / (torch.sqrt(v / c2) + self.eps) for m, v in zip(mu, nu)]
        return upd, {"count": count, "mu": mu, "nu": nu}


def pretrain_loss_cotangent(predict: List[torch.Tensor], target: List[torch.Tensor], full_det: bool = False):
    """Loss of pretrain.py:70-89 and d loss / d conj(predict) scaled so that d loss = sum Re(conj(cot) d predict):
    loss = mean_blocks mean |target[:, None] - predict|^2."""
    if full_det:                                         # block-diagonal target (pretrain.py:80-86)
        B, na, nb = target[0].shape[0], target[0].shape[1], target[1].shape[1]
        t = torch.zeros(B, na + nb, na + nb, dtype=target[0].dtype, device=target[0].device)
        t[:, :na, :na] = target[0]
        t[:, na:, na:] = target[1]
        target = [t]
    losses, cots = [], []
    nblk = len(predict)
    for tar, pre in zip(target, predict):
        diff = pre - tar[:, None, ...]
        losses.append((diff.abs() ** 2).mean())
        cots.append(2.0 * diff / (diff.numel() * nblk))
    return torch.stack(losses).mean(), cots


def make_pretrain_step(batch_orbitals, batch_network, latvec, optimizer, full_det: bool = False):
    """pretrain.py:43-111.  ``batch_orbitals`` / ``batch_network``: the ``.apply`` of deepsolid_b200 networks built with
    method_name 'eval_mats' / 'eval_slogdet' on the same HotPath.  ``key``: int seed of the device RNG (or (xi, u))."""
    getter = getattr(batch_orbitals, "hotpath", None)
    if getter is None:
        raise TypeError("batch_orbitals must be the .apply of a deepsolid_b200 network (eval_mats)")
    mh = None

    def pretrain_step(data, target, params, state, key):
        nonlocal mh
        hp = getter()
        hp.set_params(params)
        x = torch.as_tensor(data, dtype=torch.float64).to(hp.tdev)
        target = [torch.as_tensor(t).to(hp.tdev).to(torch.complex128) for t in target]
        predict = hp.orbitals(x)
        loss_val, cots = pretrain_loss_cotangent(predict, target, full_det)
        loss_val = dist.pmean(loss_val.reshape(1))[0]
        grads = hp.orbitals_vjp(x, cots)
        n_layers, bias_orb = len(params["single"]), "b" in params["orbital"][0]
        last = len(params["double"]) == n_layers         # use_last_layer pytree
        g = [dist.pmean(t) for t in flatten_params(grads)]
        updates, state = optimizer.update(unflatten_params(g, n_layers, bias_orb, last), state, params)
        leaves = [torch.as_tensor(p).to(hp.tdev) + u for p, u in zip(flatten_params(params), updates)]
        params = unflatten_params(leaves, n_layers, bias_orb, last)
        if mh is None:                                   # one move of the default width (qmc.mh_update stddev=0.02)
            mh = qmc.make_mcmc_step(batch_network, x.shape[0], latvec, steps=1)
        logprob = 2.0 * batch_network(params, x)
        data, pmove = mh(params, x, key, 0.02)
        return data, params, state, loss_val, logprob, pmove * x.shape[0]

    return pretrain_step


class PlaneWaveSCF:
    """Free-electron stand-in for hf.SCF: orbital o of spin s is exp(i k_o . r) / sqrt(V) on the occupied k-points of
    ``klist`` (what HF gives for the homogeneous electron gas).  Only ``eval_orb_mat`` / ``eval_slogdet`` of hf.py."""

    def __init__(self, simulation_cell, klist):
        self.klist = [np.asarray(k, dtype=np.float64).reshape(-1, 3) for k in klist]
        self.norm = 1.0 / math.sqrt(abs(float(np.linalg.det(np.asarray(simulation_cell.a)))))
        self.nelec = tuple(simulation_cell.nelec)

    def eval_orb_mat(self, coord):
        coord = np.asarray(coord, dtype=np.float64)
        out, o = [], 0
        for k, ns in zip(self.klist, self.nelec):
            x = coord[:, o:o + ns, :]
            out.append(self.norm * np.exp(1j * np.einsum("bic,oc->bio", x, k)))
            o += ns
        return out

    def eval_slogdet(self, coord):
        dets = [np.linalg.slogdet(m) for m in self.eval_orb_mat(coord) if m.shape[-1] > 0]
        phase = np.prod([d[0] for d in dets], axis=0)
        return phase, np.sum([d[1] for d in dets], axis=0)


def pretrain_hartree_fock(params, data, batch_network, batch_orbitals, sharded_key, cell, scf_approx,
                          full_det: bool = False, iterations: int = 1000, learning_rate: float = 5e-3,
                          log: Optional[Callable] = None):
    """pretrain.py:114-168: ``iterations`` Adam steps towards the orbitals of ``scf_approx`` on walkers that follow the
    network density.  ``data``: (B, 3N).  Returns (params, data)."""
    optimizer = Adam(learning_rate)
    state = optimizer.init(params)
    step = make_pretrain_step(batch_orbitals, batch_network, cell.lattice_vectors(), optimizer, full_det=full_det)
    key = int(sharded_key)
    nelec = [n for n in cell.nelec if n > 0]
    for t in range(iterations):
        coord = np.asarray(torch.as_tensor(data).detach().cpu(), dtype=np.float64).reshape(-1, cell.nelectron, 3)
        target = [torch.as_tensor(m) for m, ne in zip(scf_approx.eval_orb_mat(coord), cell.nelec) if ne > 0]
        key += 1
        data, params, state, loss, logprob, num_accepts = step(data, target, params, state, key)
        if log is not None:
            slog_t = sum(2.0 * torch.linalg.slogdet(m)[1] for m in target)
            log("Pretrain iter %05d: Loss=%03.6f, pmove=%0.2f, Norm of Net prob=%03.4f, Norm of HF prob=%03.4f"
                % (t, float(loss), float(num_accepts) / data.shape[0], float(logprob.mean()), float(slog_t.mean())))
    del nelec
    return params, data

"""KFAC curvature blocks on top of ds_kfac_factors (second half).

The reference's optimiser (DeepSolid/utils/kfac_ferminet_alpha) traces the jaxpr of ``total_energy`` to find the tagged
layers; a CUDA callable cannot be traced, so the drop-in is "produce the statistics the tracer would have produced, then
run the same block algebra on them":

* ``curvature_estimate``  - per tagged dense layer (network.py:443 ``register_repeated_dense``) the batch factors
  ``inputs_factor = E[(x,1)(x,1)^T]``, ``outputs_factor = E[Re(dy^H dy)]`` (curvature_blocks.py:262-281 through
  ``RepeatedDenseBlock``, DeepSolid/curvature_tags_and_blocks.py:142-156) with ``dy = sqrt2 (ga - i gp)`` from the
  fisher_exact tangent on conj(log psi) (train.py:128-133; estimator.py:284-320; loss_functions.py:529-537; vjp_rc.py);
  for the untagged envelope leaves the NaiveDiagonal statistic ``dw dw / batch`` (curvature_blocks.py:111-133).
* ``Optimizer`` - moving averages (utils.py:265-294), pi-adjusted inverses (utils.py:155-218), the preconditioned
  step with its norm constraint (optimizer.py:400-470, 572-598) for ``estimation_mode='fisher_exact'`` and momentum 0
  (base_config.py:62-75).

All factor arithmetic is torch fp64 on the device of the HotPath; the batch statistics come from the CUDA library.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch

from . import dist
from .hotpath import flatten_params, unflatten_params

_KINDS = ("single", "double", "orbital")


def curvature_estimate(hp, params, data, sync: bool = True) -> Dict[str, list]:
    """Batch curvature statistics of every block for walkers ``data`` (B, 3N) at ``params``.

    Returns ``{'single': [...], 'double': [...], 'orbital': [...]}`` with ``inputs_factor``, ``outputs_factor``,
    ``extra_scale`` per layer and ``'envelope': [{'pi', 'sigma'}, ...]`` complex diagonal factors.  With ``sync`` the
    per-rank factors are averaged over ranks (the reference averages per-device factors, utils.py:293-294)."""
    hp.set_params(params)
    raw = hp.kfac_factors(data)
    B = raw["batch"]
    out: Dict[str, list] = {k: [] for k in _KINDS}
    out["envelope"] = []
    red = dist.pmean if sync else (lambda t: t)
    for kind in _KINDS:
        for i, blk in enumerate(raw[kind]):
            has_bias = "b" in params[kind][i]
            rows = max(int(blk["rows"]), 1)
            a = blk["a"] / rows
            if not has_bias:
                a = a[:-1, :-1]
            g = 2.0 * blk["g"] / rows
            out[kind].append({"inputs_factor": red(a.contiguous()), "outputs_factor": red(g.contiguous()),
                              "extra_scale": int(blk["rows"]) // max(B, 1), "has_bias": has_bias})
    s2 = math.sqrt(2.0)
    for ga, gp in zip(raw["envelope_abs"], raw["envelope_phase"]):
        blk = {}
        for leaf in ("pi", "sigma"):
            dw = s2 * torch.complex(ga[leaf], -gp[leaf])
            blk[leaf] = red(dw * dw / max(B, 1))
        out["envelope"].append(blk)
    return out


class WeightedMovingAverage:
    """utils.py:265-294."""

    def __init__(self, like: torch.Tensor):
        self.weight = 0.0
        self.array = torch.zeros_like(like)

    def update(self, value: torch.Tensor, old_weight_multiplier: float, new_weight: float) -> None:
        self.weight = old_weight_multiplier * self.weight + new_weight
        self.array = old_weight_multiplier * self.array + new_weight * value

    @property
    def value(self) -> torch.Tensor:
        return self.array / self.weight


def psd_inv_cholesky(matrix: torch.Tensor, damping) -> torch.Tensor:
    """utils.py:130-134."""
    n = matrix.shape[0]
    eye = torch.eye(n, dtype=matrix.dtype, device=matrix.device)
    return torch.cholesky_solve(eye, torch.linalg.cholesky(matrix + damping * eye))


def pi_adjusted_inverse(factor_0: torch.Tensor, factor_1: torch.Tensor, damping: float):
    """utils.py:155-218 (the traces are already identical on every rank: the factors were averaged)."""
    norm_0, norm_1 = float(torch.trace(factor_0)), float(torch.trace(factor_1))
    scale = norm_0 * norm_1
    if not scale > 0.0:
        r = 1.0 / math.sqrt(damping)
        return (torch.eye(factor_0.shape[0], dtype=factor_0.dtype, device=factor_0.device) * r,
                torch.eye(factor_1.shape[0], dtype=factor_1.dtype, device=factor_1.device) * r)
    n0, n1 = factor_0.shape[0], factor_1.shape[0]
    if n0 == 1 and n1 == 1:
        v = torch.ones_like(factor_0) / math.sqrt(scale)
        return v, v
    if n0 == 1:
        return torch.full_like(factor_0, scale), psd_inv_cholesky(factor_1 / norm_1, damping / norm_1)
    if n1 == 1:
        return psd_inv_cholesky(factor_0 / norm_0, damping / norm_0), torch.full_like(factor_1, scale)
    d0 = math.sqrt(damping * n1 / (scale * n0))
    d1 = math.sqrt(damping * n0 / (scale * n1))
    return (psd_inv_cholesky(factor_0 / norm_0, d0) / math.sqrt(scale),
            psd_inv_cholesky(factor_1 / norm_1, d1) / math.sqrt(scale))


class Optimizer:
    """The reference's ``kfac_optim.Optimizer`` (process.py:209-222) reduced to the configuration DeepSolid runs:
    fisher_exact curvature, fixed damping, regular momentum (0 in base_config.py:67), norm constraint; adaptive damping /
    momentum are "not currently available" in the reference either (base_config.py:69).

    ``value_and_grad(params, data) -> ((loss, aux), grads)`` is ``train.make_loss(...).value_and_grad``, which returns
    the gradient of THIS rank's walkers; ``step`` averages it (and the norm-constraint scalar) over ranks, as the
    reference does (kfac_ferminet_alpha optimizer.py:423, :593), so walker-sharded replicas stay identical.
    ``hp`` is the HotPath the loss evaluates."""

    def __init__(self, value_and_grad: Callable, hp, l2_reg: float = 0.0, norm_constraint: Optional[float] = 1e-3,
                 curvature_ema: float = 0.95, inverse_update_period: int = 1, cov_update_every: int = 1):
        self.value_and_grad = value_and_grad
        self.hp = hp
        self.l2_reg = float(l2_reg)
        self.norm_constraint = norm_constraint
        self.curvature_ema = float(curvature_ema)
        self.inverse_update_period = int(inverse_update_period)
        self.cov_update_every = int(cov_update_every)
        self.step_counter = 0
        self.velocities = None          # last update (list of leaves), optimizer.py:617-627
        self.blocks: Optional[Dict[str, list]] = None

    # -- curvature state ---------------------------------------------------
    def _init_blocks(self, est) -> None:
        self.blocks = {k: [] for k in _KINDS}
        self.blocks["envelope"] = []
        for kind in _KINDS:
            for b in est[kind]:
                self.blocks[kind].append({"inputs": WeightedMovingAverage(b["inputs_factor"]),
                                          "outputs": WeightedMovingAverage(b["outputs_factor"]),
                                          "extra_scale": b["extra_scale"], "has_bias": b["has_bias"],
                                          "inputs_inv": None, "outputs_inv": None})
        for b in est["envelope"]:
            self.blocks["envelope"].append({leaf: WeightedMovingAverage(b[leaf]) for leaf in ("pi", "sigma")})

    def update_curvature_matrix_estimate(self, est, ema_old: float, ema_new: float) -> None:
        """estimator.py:239-320 for one fisher_exact index (total_num_indices = 1)."""
        if self.blocks is None:
            self._init_blocks(est)
        for kind in _KINDS:
            for blk, b in zip(self.blocks[kind], est[kind]):
                blk["inputs"].update(b["inputs_factor"], ema_old, ema_new)
                blk["outputs"].update(b["outputs_factor"], ema_old, ema_new)
        for blk, b in zip(self.blocks["envelope"], est["envelope"]):
            for leaf in ("pi", "sigma"):
                blk[leaf].update(b[leaf], ema_old, ema_new)

    def update_curvature_estimate_inverse(self, damping: float) -> None:
        """curvature_blocks.py:193-213: damping / extra_scale, pi-adjusted."""
        for kind in _KINDS:
            for blk in self.blocks[kind]:
                blk["inputs_inv"], blk["outputs_inv"] = pi_adjusted_inverse(
                    blk["inputs"].value, blk["outputs"].value, damping / blk["extra_scale"])

    def multiply_inverse(self, grads, damping: float):
        """estimator.multiply_matpower(grads, -1) (estimator.py:217-237; curvature_blocks.py:215-247, 142-154)."""
        out = {k: [] for k in _KINDS}
        out["envelope"] = []
        for kind in _KINDS:
            for blk, g in zip(self.blocks[kind], grads[kind]):
                w = g["w"]
                vec = torch.cat([w, g["b"][None]], dim=0) if blk["has_bias"] else w
                res = blk["inputs_inv"] @ vec @ blk["outputs_inv"] / blk["extra_scale"]
                out[kind].append({"w": res[:-1], "b": res[-1]} if blk["has_bias"] else {"w": res})
        for blk, g in zip(self.blocks["envelope"], grads["envelope"]):
            # the reference divides by a complex diagonal (dw*dw of a complex tangent); the real part keeps the
            # parameters real
            out["envelope"].append({leaf: (g[leaf] / (blk[leaf].value + damping)).real for leaf in ("pi", "sigma")})
        return out

    # -- one optimisation step --------------------------------------------
    def step(self, params, data, learning_rate: float, damping: float, momentum: float = 0.0):
        """optimizer.py:400-470: gradients, curvature EMA, (periodic) inverses, preconditioned direction with the norm
        constraint, params + delta.  -> (new params, stats dict with loss / aux / coefficient)."""
        (loss, aux), grads = self.value_and_grad(params, data)
        dev = self.hp.tdev
        to_dev = lambda t: torch.as_tensor(t).to(dev)
        leaves_p = [to_dev(t) for t in flatten_params(params)]
        leaves_g = [dist.pmean(to_dev(t)) for t in flatten_params(grads)]        # optimizer.py:423
        if self.l2_reg:
            leaves_g = [g + self.l2_reg * p for g, p in zip(leaves_g, leaves_p)]
        n_layers = len(params["single"])
        bias_orb = "b" in params["orbital"][0]
        grads = unflatten_params(leaves_g, n_layers, bias_orb)
        if self.step_counter % self.cov_update_every == 0:
            est = curvature_estimate(self.hp, params, data)
            self.update_curvature_matrix_estimate(est, self.curvature_ema, 1.0)
        if self.step_counter % self.inverse_update_period == 0 or self.blocks["single"][0]["inputs_inv"] is None:
            self.update_curvature_estimate_inverse(damping)
        pre = self.multiply_inverse(grads, damping)
        leaves_pre = flatten_params(pre)
        coefficient = 1.0
        if self.norm_constraint is not None:
            sq = sum((a * b).sum() for a, b in zip(leaves_pre, leaves_g)) * learning_rate ** 2
            sq = float(dist.pmean(sq))                                               # optimizer.py:593
            coefficient = min(math.sqrt(self.norm_constraint / sq), 1.0) if sq > 0.0 else 1.0
        # optimizer.py:444-463, 617-627: vectors = (preconditioned gradient, velocities), coefficients = (-lr, momentum);
        # delta = sum of the two, velocities <- delta; the first step treats the momentum as 0 (optimizer.py:562-564)
        mom = float(momentum) if (self.velocities is not None and self.step_counter > 0) else 0.0
        delta = [-learning_rate * coefficient * d for d in leaves_pre]
        if mom != 0.0:
            delta = [d + mom * v for d, v in zip(delta, self.velocities)]
        self.velocities = delta
        new_leaves = [p + d for p, d in zip(leaves_p, delta)]
        self.step_counter += 1
        return unflatten_params(new_leaves, n_layers, bias_orb), {"loss": loss, "aux": aux, "coefficient": coefficient,
                                                                  "momentum": mom}

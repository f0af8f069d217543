"""Device context of the hot path: one ``HotPath`` per (simulation cell, network
configuration, CUDA device).  It owns the C-ABI context, uploads the parameter
pytree when it changes, and exposes batched torch-tensor entry points that the
reference-shaped closures in network.py / hamiltonian.py / qmc.py / train.py call.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ewald_tables import build_ewald_tables

LAP_MODES = {"for": 0, "hessian": 1, "dim_batch": 2, "partition": 3}


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_double_p)


def flatten_params(params) -> list:
    """Leaf order of ds_set_params (include/deepsolid_b200.h)."""
    leaves = []
    for layer in params["single"]:
        leaves += [layer["w"], layer["b"]]
    for layer in params["double"]:
        leaves += [layer["w"], layer["b"]]
    for orb in params["orbital"]:
        leaves.append(orb["w"])
        if "b" in orb:
            leaves.append(orb["b"])
    for env in params["envelope"]:
        leaves += [env["pi"], env["sigma"]]
    return leaves


def unflatten_params(leaves, n_layers: int, bias_orbitals: bool = False, use_last_layer: bool = False) -> dict:
    """Inverse of flatten_params."""
    it = iter(leaves)
    single = [{"w": next(it), "b": next(it)} for _ in range(n_layers)]
    double = [{"w": next(it), "b": next(it)} for _ in range(n_layers if use_last_layer else n_layers - 1)]
    orbital = [({"w": next(it), "b": next(it)} if bias_orbitals else {"w": next(it)}) for _ in range(2)]
    envelope = [{"pi": next(it), "sigma": next(it)} for _ in range(2)]
    return {"single": single, "double": double, "orbital": orbital, "envelope": envelope}


class HotPath:
    def __init__(self, simulation_cell, klist, hidden_dims=((256, 32),) * 3, determinants: int = 8,
                 device: Optional[int] = None, distance_type: str = "nu", envelope_type: str = "isotropic",
                 bias_orbitals: bool = False, full_det: bool = False, use_last_layer: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("deepsolid_b200 needs a CUDA device: the local-energy hot path has no CPU fallback")
        self.lib = _lib.load()
        self.cell = simulation_cell
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        prim = simulation_cell.original_cell
        hidden_dims = tuple(tuple(h) for h in hidden_dims)
        if len({h[0] for h in hidden_dims}) != 1 or len({h[1] for h in hidden_dims}) != 1:
            raise ValueError("the CUDA hot path needs equal widths in every layer of hidden_dims")
        self.hidden_dims = hidden_dims
        self.determinants = int(determinants)
        if distance_type not in ("nu", "tri"):
            raise ValueError("Unrecognized distance function.")
        self.distance_type = distance_type
        envs = {"isotropic": 0, "diagonal": 1, "full": 2}
        if envelope_type not in envs:
            raise ValueError(f"envelope_type={envelope_type!r} is not implemented in the CUDA hot path")
        if envelope_type != "isotropic" and distance_type != "nu":
            raise ValueError("diagonal / full envelopes need distance_type='nu' (3-component relative vectors)")
        self.envelope_type = envelope_type
        self.bias_orbitals = bool(bias_orbitals)
        self.full_det = bool(full_det)
        self.use_last_layer = bool(use_last_layer)
        if self.use_last_layer and len(hidden_dims) > 3:
            raise ValueError("use_last_layer=True is implemented for at most 3 layers")
        self.n_up, self.n_dn = simulation_cell.nelec
        self.nelec = self.n_up + self.n_dn
        tb = build_ewald_tables(simulation_cell)
        self.tables = tb
        ne = self.nelec
        keep = [_f64(prim.a), _f64(simulation_cell.a), _f64(prim.AV), _f64(prim.BV), _f64(simulation_cell.AV),
                _f64(simulation_cell.BV), _f64(prim.atom_coords()), _f64(tb.atom_coords), _f64(tb.atom_charges),
                _f64(klist[0]).reshape(-1, 3), _f64(klist[1]).reshape(-1, 3), _f64(tb.mi_shifts),
                _f64(tb.lattice_displacements), _f64(tb.gpoints), _f64(tb.gweight),
                _f64(tb.ion_exp.real), _f64(tb.ion_exp.imag)]
        for m in keep[2:6]:
            if m.shape != (3, 3):
                raise ValueError("only sym_type='minimal' (3 reciprocal rows) is implemented")
        if keep[9].shape[0] != self.n_up or keep[10].shape[0] != self.n_dn:
            raise ValueError("klist must hold one k-point per occupied orbital of each spin")
        sd = _lib.SystemDesc(
            n_up=self.n_up, n_dn=self.n_dn, n_atoms_prim=prim.natm, n_atoms_sim=len(tb.atom_charges),
            prim_latvec=_ptr(keep[0]), sim_latvec=_ptr(keep[1]), prim_AV=_ptr(keep[2]), prim_BV=_ptr(keep[3]),
            sim_AV=_ptr(keep[4]), sim_BV=_ptr(keep[5]), prim_atoms=_ptr(keep[6]), sim_atoms=_ptr(keep[7]),
            sim_charges=_ptr(keep[8]), klist_up=_ptr(keep[9]), klist_dn=_ptr(keep[10]),
            dist_kind=int(tb.dist_kind), mi_shifts=_ptr(keep[11]), lattice_displacements=_ptr(keep[12]),
            alpha=float(tb.alpha), n_g=int(len(tb.gweight)), gpoints=_ptr(keep[13]), gweight=_ptr(keep[14]),
            ion_exp_re=_ptr(keep[15]), ion_exp_im=_ptr(keep[16]),
            ee_const=float(tb.ee_const(ne)), ei_const=float(tb.ei_const(ne)), ii_total=float(tb.ii_total))
        nd = _lib.NetDesc(n_layers=len(hidden_dims), hidden_one=hidden_dims[0][0], hidden_two=hidden_dims[0][1],
                          n_det=self.determinants, distance_type=1 if distance_type == "tri" else 0,
                          envelope_type=envs[envelope_type], bias_orbitals=1 if bias_orbitals else 0,
                          full_det=1 if full_det else 0, use_last_layer=1 if use_last_layer else 0)
        h = C.c_void_p()
        _lib.check(self.lib.ds_ctx_create(C.byref(sd), C.byref(nd), self.device, C.byref(h)))
        self.h = h
        self._param_key = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ds_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------
    def set_params(self, params, force: bool = False) -> None:
        """Upload the reference parameter pytree (network.py:135-184) if it changed.

        "Changed" is detected from (data_ptr, torch version counter, shape) of every leaf: an in-place edit made
        through a numpy view or ``tensor.data`` does not bump the version counter -- pass ``force=True`` after one.
        ``ds_set_params`` copies on the legacy default stream and synchronises it; device leaves written on the
        current torch stream are made visible by synchronising that stream first."""
        leaves = flatten_params(params)
        tens = []
        for lf in leaves:
            t = lf if isinstance(lf, torch.Tensor) else torch.as_tensor(np.asarray(lf))
            if t.dtype != torch.float64 or not t.is_contiguous():
                t = t.to(torch.float64).contiguous()
            tens.append(t)
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tens)
        if key == self._param_key and not force:
            return
        if any(t.is_cuda for t in tens):
            torch.cuda.current_stream(self.tdev).synchronize()
        n = len(tens)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tens])
        sizes = (C.c_int64 * n)(*[t.numel() for t in tens])
        _lib.check(self.lib.ds_set_params(self.h, ptrs, sizes, n))
        self._param_key = key
        self._keep_params = tens

    def set_workspace_limit(self, nbytes: int) -> None:
        _lib.check(self.lib.ds_set_workspace_limit(self.h, int(nbytes)))

    # ------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def _prep(self, x) -> Tuple[torch.Tensor, bool, bool]:
        """-> (2-D float64 contiguous tensor, was_1d, on_device)."""
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        one = t.dim() == 1
        if one:
            t = t[None]
        if t.dim() != 2 or t.shape[1] != 3 * self.nelec:
            raise ValueError(f"walkers must have shape (batch, {3 * self.nelec}); got {tuple(t.shape)}")
        t = t.to(torch.float64).contiguous()
        on_dev = t.is_cuda
        if on_dev and t.device != self.tdev:
            raise ValueError(f"walkers live on {t.device}, context on {self.tdev}")
        return t, one, on_dev

    def logpsi(self, x):
        """(log|psi|, angle) for a batch; device tensors stay on the device, host tensors go
        through the C-ABI host entry point (copies included)."""
        t, one, on_dev = self._prep(x)
        B = t.shape[0]
        if on_dev:
            la = torch.empty(B, dtype=torch.float64, device=self.tdev)
            ph = torch.empty_like(la)
            _lib.check(self.lib.ds_logpsi(self.h, t.data_ptr(), B, la.data_ptr(), ph.data_ptr(), self._stream()))
        else:
            la = torch.empty(B, dtype=torch.float64)
            ph = torch.empty_like(la)
            _lib.check(self.lib.ds_logpsi_host(self.h, t.data_ptr(), B, la.data_ptr(), ph.data_ptr()))
        return (la[0], ph[0]) if one else (la, ph)

    def logpsi_vjp(self, x, cot_abs, cot_phase):
        """Parameter gradient pytree of sum_b cot_abs[b] log|psi_b| + cot_phase[b] angle(psi_b)
        (the reverse sweep behind train.make_loss's custom JVP, train.py:129-137).  Device tensors in,
        device tensors out, same pytree structure as the parameters last given to ``set_params``."""
        t, one, on_dev = self._prep(x)
        td = t if on_dev else t.to(self.tdev)
        B = td.shape[0]
        ca = torch.as_tensor(cot_abs, dtype=torch.float64).reshape(-1).to(self.tdev).contiguous()
        cp = torch.as_tensor(cot_phase, dtype=torch.float64).reshape(-1).to(self.tdev).contiguous()
        if ca.numel() != B or cp.numel() != B:
            raise ValueError("cotangents must have one entry per walker")
        if self._param_key is None:
            raise ValueError("parameters have not been set")
        outs = [torch.empty(tp.shape, dtype=torch.float64, device=self.tdev) for tp in self._keep_params]
        n = len(outs)
        ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        sizes = (C.c_int64 * n)(*[o.numel() for o in outs])
        _lib.check(self.lib.ds_logpsi_vjp(self.h, td.data_ptr(), B, ca.data_ptr(), cp.data_ptr(), ptrs, sizes, n,
                                          self._stream()))
        return unflatten_params(outs, len(self.hidden_dims), self.bias_orbitals, self.use_last_layer)

    def orbitals_vjp(self, x, cot_mats):
        """Parameter gradient pytree of sum Re(conj(cot) * M) over the orbital matrices of ``orbitals(x)``
        (eval_mats, network.py:601-602; pretrain.py:70-89).  ``cot_mats``: list of two complex tensors
        (B, D, n_s, n_s), the cotangents of the spin-up / spin-down matrices."""
        t, one, on_dev = self._prep(x)
        td = t if on_dev else t.to(self.tdev)
        B = td.shape[0]
        D = self.determinants
        parts = []
        for cm, ns in zip(cot_mats, ((self.nelec,) if self.full_det else (self.n_up, self.n_dn))):
            cm = torch.as_tensor(cm).to(self.tdev).to(torch.complex128).reshape(B, D * ns * ns)
            parts.append(torch.view_as_real(cm).reshape(B, -1))
        cot = torch.cat(parts, dim=1).contiguous()
        if cot.shape[1] != int(self.lib.ds_orbitals_size(self.h)):
            raise ValueError("cotangent does not have the shape of the orbital matrices")
        if self._param_key is None:
            raise ValueError("parameters have not been set")
        outs = [torch.empty(tp.shape, dtype=torch.float64, device=self.tdev) for tp in self._keep_params]
        n = len(outs)
        ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        sizes = (C.c_int64 * n)(*[o.numel() for o in outs])
        _lib.check(self.lib.ds_orbitals_vjp(self.h, td.data_ptr(), B, cot.data_ptr(), ptrs, sizes, n, self._stream()))
        return unflatten_params(outs, len(self.hidden_dims), self.bias_orbitals, self.use_last_layer)

    def rho_q(self, x, qvecs, mode: int = 0):
        """Complex (B, nq) plane-wave sums over the electrons: mode 0 sum_i exp(i q.x_i), mode 1 exp(i sum_i q.x_i)
        (estimator.py:27-31, 68-70)."""
        t, one, on_dev = self._prep(x)
        td = t if on_dev else t.to(self.tdev)
        q = torch.as_tensor(np.asarray(qvecs, dtype=np.float64).reshape(-1, 3)).to(self.tdev).contiguous()
        B, nq = td.shape[0], q.shape[0]
        out = torch.empty(B, nq, 2, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.ds_rho_q(self.h, td.data_ptr(), B, q.data_ptr(), nq, int(mode), out.data_ptr(), self._stream()))
        res = torch.view_as_complex(out)
        if not on_dev:
            res = res.cpu()
        return res[0] if one else res

    def kfac_factors(self, x):
        """Raw Kronecker-factor sums of every tagged dense layer over the walkers ``x`` (ds_kfac_factors): what the
        reference's KFAC estimator extracts from total_energy_jvp (train.py:128-133; kfac_ferminet_alpha/
        estimator.py:284-320, curvature_blocks.py:262-281; curvature_tags_and_blocks.py:142-156).

        Returns a dict with, per tagged layer of ``single``, ``double``, ``orbital``: ``a`` = sum_rows (x,1)(x,1)^T,
        ``g`` = sum_rows ga ga^T + gp gp^T and ``rows``; plus ``envelope_abs`` / ``envelope_phase``: gradients of
        sum_w log|psi_w| / sum_w angle(psi_w) w.r.t. the (untagged) envelope leaves.  deepsolid_b200.kfac turns these
        into the reference's factors."""
        t, one, on_dev = self._prep(x)
        td = t if on_dev else t.to(self.tdev)
        B = td.shape[0]
        if self._param_key is None:
            raise ValueError("parameters have not been set")
        L = len(self.hidden_dims)
        leaves = self._keep_params
        kinds, nin, nout, rows = [], [], [], []
        li = 0
        for l in range(L):
            kinds.append(("single", l)); nin.append(leaves[li].shape[0]); nout.append(leaves[li].shape[1])
            rows.append(B * self.nelec); li += 2
        for l in range(L if self.use_last_layer else L - 1):
            kinds.append(("double", l)); nin.append(leaves[li].shape[0]); nout.append(leaves[li].shape[1])
            rows.append(B * self.nelec * self.nelec); li += 2
        for s, ns in enumerate((self.n_up, self.n_dn)):
            kinds.append(("orbital", s)); nin.append(leaves[li].shape[0]); nout.append(leaves[li].shape[1])
            rows.append(B * ns); li += 2 if self.bias_orbitals else 1
        env_leaves = leaves[li:li + 4]
        a = [torch.empty(n + 1, n + 1, dtype=torch.float64, device=self.tdev) for n in nin]
        g = [torch.empty(n, n, dtype=torch.float64, device=self.tdev) for n in nout]
        ea = [torch.empty(tp.shape, dtype=torch.float64, device=self.tdev) for tp in env_leaves]
        ep = [torch.empty(tp.shape, dtype=torch.float64, device=self.tdev) for tp in env_leaves]
        n = len(a)

        def arr(ts):
            return ((C.c_void_p * len(ts))(*[o.data_ptr() for o in ts]), (C.c_int64 * len(ts))(*[o.numel() for o in ts]))

        ap, asz = arr(a)
        gp, gsz = arr(g)
        eap, esz = arr(ea)
        epp, _ = arr(ep)
        _lib.check(self.lib.ds_kfac_factors(self.h, td.data_ptr(), B, ap, asz, gp, gsz, n, eap, epp, esz, 4,
                                            self._stream()))
        out = {"single": [], "double": [], "orbital": []}
        for (kind, _), ai, gi, r in zip(kinds, a, g, rows):
            out[kind].append({"a": ai, "g": gi, "rows": r})
        out["envelope_abs"] = [{"pi": ea[0], "sigma": ea[1]}, {"pi": ea[2], "sigma": ea[3]}]
        out["envelope_phase"] = [{"pi": ep[0], "sigma": ep[1]}, {"pi": ep[2], "sigma": ep[3]}]
        out["batch"] = B
        return out

    def logpsi_grad_x(self, x, want_phase_grad: bool = False):
        """(log|psi|, phase, d log|psi|/dx [, d phase/dx]) on the device: jax.value_and_grad of the slog network
        w.r.t. the walker (qmc.py:325), from the first-derivative half of the forward-Laplacian sweep."""
        t, one, on_dev = self._prep(x)
        td = t if on_dev else t.to(self.tdev)
        B = td.shape[0]
        la = torch.empty(B, dtype=torch.float64, device=self.tdev)
        ph = torch.empty_like(la)
        ga = torch.empty(B, 3 * self.nelec, dtype=torch.float64, device=self.tdev)
        gp = torch.empty_like(ga) if want_phase_grad else None
        _lib.check(self.lib.ds_logpsi_grad_x(self.h, td.data_ptr(), B, la.data_ptr(), ph.data_ptr(), ga.data_ptr(),
                                             gp.data_ptr() if gp is not None else None, self._stream()))
        out = (la, ph, ga) + ((gp,) if want_phase_grad else ())
        if not on_dev:
            out = tuple(o.cpu() for o in out)
        return tuple(o[0] for o in out) if one else out

    def orbitals(self, x):
        t, one, on_dev = self._prep(x)
        B = t.shape[0]
        td = t if on_dev else t.to(self.tdev)
        per = int(self.lib.ds_orbitals_size(self.h))
        out = torch.empty(B, per, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.ds_orbitals(self.h, td.data_ptr(), B, out.data_ptr(), self._stream()))
        D = self.determinants
        mats, o = [], 0
        for ns in ((self.nelec,) if self.full_det else (self.n_up, self.n_dn)):
            n = D * ns * ns * 2
            m = torch.view_as_complex(out[:, o:o + n].reshape(B, D, ns, ns, 2).contiguous())
            mats.append(m if on_dev else m.cpu())
            o += n
        return [m[0] for m in mats] if one else mats

    def local_energy(self, x, mode: str = "for", partition_number: int = 3):
        """(kinetic complex128 [B], ewald float64 [B])."""
        if mode not in LAP_MODES:
            raise ValueError("Unrecognized laplacian evaluation mode.")
        t, one, on_dev = self._prep(x)
        B = t.shape[0]
        dev = self.tdev if on_dev else torch.device("cpu")
        out = torch.empty(3, B, dtype=torch.float64, device=dev)
        if on_dev:
            _lib.check(self.lib.ds_local_energy(self.h, t.data_ptr(), B, LAP_MODES[mode], int(partition_number),
                                                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                                self._stream()))
        else:
            _lib.check(self.lib.ds_local_energy_host(self.h, t.data_ptr(), B, LAP_MODES[mode], int(partition_number),
                                                     out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()))
        ke = torch.complex(out[0], out[1])
        ew = out[2]
        return (ke[0], ew[0]) if one else (ke, ew)

    def ewald(self, x):
        t, one, on_dev = self._prep(x)
        td = t if on_dev else t.to(self.tdev)
        B = td.shape[0]
        out = torch.empty(2, B, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.ds_ewald(self.h, td.data_ptr(), B, out[0].data_ptr(), out[1].data_ptr(), self._stream()))
        ii = torch.full((B,), float(self.lib.ds_ewald_ii(self.h)), dtype=torch.float64, device=self.tdev)
        res = (out[0], out[1], ii)
        if not on_dev:
            res = tuple(r.cpu() for r in res)
        return tuple(r[0] for r in res) if one else res

    def mcmc(self, data, steps: int, width: float, seed: int = 0, xi=None, u=None, return_masks: bool = False,
             one_electron: bool = False):
        """In-place-free Metropolis sweep: returns (new_data, n_accept tensor[1], masks or None).
        ``one_electron``: steps * N single-electron moves (qmc.py:227-287, 355-358); noise shapes are then
        xi (steps*N, batch, 3), u (steps*N, batch)."""
        t, one, on_dev = self._prep(data)
        if one:
            raise ValueError("mcmc_step needs batched walkers (batch, 3N)")
        B = t.shape[0]
        if one_electron:
            ns = steps * self.nelec
            x = t.clone() if on_dev else t.to(self.tdev)
            nacc = torch.zeros(1, dtype=torch.float64, device=self.tdev)
            masks = torch.empty(ns, B, dtype=torch.uint8, device=self.tdev) if return_masks else None
            xi_d = xi.to(self.tdev, torch.float64).contiguous() if xi is not None else None
            u_d = u.to(self.tdev, torch.float64).contiguous() if u is not None else None
            if xi_d is not None and tuple(xi_d.shape) != (ns, B, 3):
                raise ValueError("xi must have shape (steps*N, batch, 3) for one-electron moves")
            if u_d is not None and tuple(u_d.shape) != (ns, B):
                raise ValueError("u must have shape (steps*N, batch) for one-electron moves")
            _lib.check(self.lib.ds_mcmc_step_one_electron(
                self.h, x.data_ptr(), B, int(steps), float(width), int(seed) & (2 ** 64 - 1),
                xi_d.data_ptr() if xi_d is not None else None, u_d.data_ptr() if u_d is not None else None,
                masks.data_ptr() if masks is not None else None, nacc.data_ptr(), self._stream()))
            if not on_dev:
                return x.cpu(), nacc.cpu(), (masks.cpu() if masks is not None else None)
            return x, nacc, masks
        if on_dev:
            x = t.clone()
            nacc = torch.zeros(1, dtype=torch.float64, device=self.tdev)
            masks = torch.empty(steps, B, dtype=torch.uint8, device=self.tdev) if return_masks else None
            xi_d = xi.to(self.tdev, torch.float64).contiguous() if xi is not None else None
            u_d = u.to(self.tdev, torch.float64).contiguous() if u is not None else None
            if xi_d is not None and tuple(xi_d.shape) != (steps, B, 3 * self.nelec):
                raise ValueError("xi must have shape (steps, batch, 3N)")
            if u_d is not None and tuple(u_d.shape) != (steps, B):
                raise ValueError("u must have shape (steps, batch)")
            _lib.check(self.lib.ds_mcmc_step(
                self.h, x.data_ptr(), B, int(steps), float(width), int(seed) & (2 ** 64 - 1),
                xi_d.data_ptr() if xi_d is not None else None, u_d.data_ptr() if u_d is not None else None,
                masks.data_ptr() if masks is not None else None, nacc.data_ptr(), self._stream()))
            return x, nacc, masks
        x = t.clone()
        nacc = torch.zeros(1, dtype=torch.float64)
        masks = torch.empty(steps, B, dtype=torch.uint8) if return_masks else None
        xi_h = xi.to(torch.float64).contiguous() if xi is not None else None
        u_h = u.to(torch.float64).contiguous() if u is not None else None
        _lib.check(self.lib.ds_mcmc_step_host(
            self.h, x.data_ptr(), B, int(steps), float(width), int(seed) & (2 ** 64 - 1),
            xi_h.data_ptr() if xi_h is not None else None, u_h.data_ptr() if u_h is not None else None,
            masks.data_ptr() if masks is not None else None, nacc.data_ptr()))
        return x, nacc, masks

    def energy_stats(self, ke: torch.Tensor, ew: torch.Tensor) -> torch.Tensor:
        """[sum Re e, sum Im e, sum |e|^2, sum Re ke, sum ew, n] on the device."""
        kr = ke.real.contiguous()
        ki = ke.imag.contiguous()
        ew = ew.contiguous()
        out = torch.empty(6, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.ds_energy_stats(self.h, kr.data_ptr(), ki.data_ptr(), ew.data_ptr(), kr.numel(),
                                            out.data_ptr(), self._stream()))
        return out

    def stats_allreduce(self, stats6: torch.Tensor, comm=None, n_accept: Optional[torch.Tensor] = None,
                        moves_per_rank: float = 0.0, global_variance: bool = False) -> torch.Tensor:
        """ds_stats_allreduce: [loss, imaginary, variance, mean Re ke, mean ewald, n_ranks, n_walkers, pmove] after
        one ncclAllReduce of the packed statistics on the current stream (``comm`` None: single rank)."""
        out = torch.empty(8, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.ds_stats_allreduce(
            self.h, comm, stats6.data_ptr(), n_accept.data_ptr() if n_accept is not None else None,
            float(moves_per_rank), 1 if global_variance else 0, out.data_ptr(), self._stream()))
        return out

    # ---- instrumentation ------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.ds_launch_count(self.h))

    def profile(self, on: bool):
        _lib.check(self.lib.ds_profile_enable(self.h, 1 if on else 0))
        _lib.check(self.lib.ds_profile_reset(self.h))

    def profile_get(self) -> Dict[str, float]:
        ms, fl, tot = C.c_double(), C.c_double(), C.c_double()
        n = C.c_int64()
        _lib.check(self.lib.ds_profile_get(self.h, C.byref(ms), C.byref(n), C.byref(fl), C.byref(tot)))
        return {"jac_ms": ms.value, "jac_launches": n.value, "jac_flops": fl.value, "total_ms": tot.value}

    def workspace_info(self) -> Dict[str, int]:
        cw, wb = C.c_int64(), C.c_int64()
        _lib.check(self.lib.ds_workspace_info(self.h, C.byref(cw), C.byref(wb)))
        return {"chunk_walkers": int(cw.value), "workspace_bytes": int(wb.value)}

    def debug_buffer(self, name: str) -> torch.Tensor:
        n = int(self.lib.ds_debug_buffer(self.h, name.encode(), None, 0))
        if n < 0:
            raise KeyError(name)
        out = torch.empty(n, dtype=torch.float64, device=self.tdev)
        if n:
            self.lib.ds_debug_buffer(self.h, name.encode(), out.data_ptr(), n)
        return out

    def debug_set(self, key: str, value: int):
        _lib.check(self.lib.ds_debug_set_int(self.h, key.encode(), int(value)))

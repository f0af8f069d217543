"""CPU accuracy study of the int8-slice (Ozaki scheme I) product used by the tcgen05 path.

A fp64 matrix is scaled per row (A) / per column (W) by a power of two and cut into S
balanced base-256 digits (int8); digit products accumulate exactly in int32 and the
diagonals s+t < S are recombined in fp64.  This script replaces the two Jacobian-row
contractions of oracle/forward_laplacian.py by that arithmetic and prints the error of
the kinetic energy against the exact fp64 evaluation, for several S.

    python scripts/ozaki_study.py [system] [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O, forward_laplacian as FL


def digits(M, S, axis):
    """Balanced base-256 digits of M scaled by the power of two covering max|M| along `axis`.
    Returns (list of S float64 digit tensors (most significant first), exponent tensor)."""
    mx = M.abs().amax(dim=axis, keepdim=True)
    e = torch.where(mx > 0, torch.floor(torch.log2(mx)) + 1, torch.zeros_like(mx))    # max < 2^e
    v = torch.round(M * torch.exp2((8 * S - 2) - e)).to(torch.int64)                  # |v| <= 2^(8S-2)
    out = []
    for _ in range(S - 1):
        d = ((v + 128) & 255) - 128
        out.append(d.to(torch.float64))
        v = (v - d) >> 8
    assert int(v.abs().max()) <= 127, int(v.abs().max())
    out.append(v.to(torch.float64))
    return out[::-1], e


def ozaki_matmul(S, SB=None, ndiag=None):
    SB = SB or S
    ndiag = ndiag or S

    def mm(A, W):
        dA, eA = digits(A, S, -1)
        dW, eW = digits(W, SB, 0)
        acc = torch.zeros(A.shape[:-1] + (W.shape[1],), dtype=torch.float64)
        for g in range(ndiag - 1, -1, -1):
            part = None
            for s in range(min(S, g + 1)):
                t = g - s
                if t >= SB:
                    continue
                p = dA[s] @ dW[t]
                part = p if part is None else part + p
            if part is not None:
                acc = acc + part * 2.0 ** (-8 * g - 12)
        return acc * torch.exp2(eA) * torch.exp2(eW)
    return mm


def main():
    system = sys.argv[1] if len(sys.argv) > 1 else "graphene8"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    sc = C.build_system(system)
    kl = C.make_klist(sc)
    P = O.params_to_torch(O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec))
    X = torch.as_tensor(C.init_walkers(sc, B, seed=3))
    FL.JAC_MATMUL = None
    la0, ph0, ke0, _ = FL.kinetic_forward_laplacian(P, X, sc, kl)
    print(f"{system}: N={sum(sc.nelec)}  |ke| max {float(ke0.abs().max()):.3f}")
    for S in (3, 4, 5, 6, 7):
        FL.JAC_MATMUL = ozaki_matmul(S)
        la, ph, ke, _ = FL.kinetic_forward_laplacian(P, X, sc, kl)
        print(f"  S={S}: {S * (S + 1) // 2:2d} int8 GEMMs  max |d ke| = {float((ke - ke0).abs().max()):.3e} Ha")
    FL.JAC_MATMUL = None


if __name__ == "__main__":
    main()

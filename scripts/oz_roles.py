"""Role clocks of oz_gemm_kernel during one local-energy pass (DS_OZ_OPT bit 32 must be set in the environment):
where the MMA thread, the TMA producer and an epilogue warp spend their time, per kernel mode.
  DS_OZ_OPT=32 python scripts/oz_roles.py [system] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepsolid_b200 import cell as C, network, hamiltonian

name = sys.argv[1] if len(sys.argv) > 1 else "graphite54"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 482
dev = torch.device("cuda", 0)
sc = C.build_system(name)
kl = C.make_klist(sc)
P = network.init_solid_fermi_net_params(888, atoms=sc.original_cell.atom_coords(), spins=sc.nelec,
                                        envelope_type="isotropic", full_det=False, determinants=8)
X = torch.as_tensor(C.init_walkers(sc, B, seed=666)).to(dev)
net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, method_name="eval_logdet", klist=kl, simulation_cell=sc, determinants=8, device=0)
hp = net.apply.hotpath()
el = hamiltonian.local_energy_seperate(net.apply, sc, mode="for")
el(P, X); torch.cuda.synchronize()
hp.debug_buffer("oz_prof")                     # clear
el(P, X); torch.cuda.synchronize()
v = hp.debug_buffer("oz_prof").cpu().numpy().reshape(8, 8)
modes = ["PLAIN", "JAC", "ORBJ", "VALUE", "LAP", "JACD"]
print(f"{name} batch {B}: clocks per tile (average over CTAs)")
print(f"{'mode':6s} {'tiles':>9s} {'tile':>8s} {'mma:drain':>10s} {'mma:stage':>10s} {'tma:free':>9s} {'epi:wait':>9s} {'epi:A':>8s} {'epi:B':>8s}")
for m, r in zip(modes, v):
    if r[7] == 0:
        continue
    t = r[7]
    print(f"{m:6s} {int(t):9d} {r[0]/t:8.0f} {r[1]/t:10.0f} {r[2]/t:10.0f} {r[3]/t:9.0f} {r[4]/t:9.0f} {r[5]/t:8.0f} {r[6]/t:8.0f}")

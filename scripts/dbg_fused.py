"""Debug helper: one local-energy pass at a given system / batch, fused vs unfused digits (prints the CUDA error, if any)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepsolid_b200 import cell as C, network, hamiltonian
name, B = sys.argv[1], int(sys.argv[2])
sc = C.build_system(name); kl = C.make_klist(sc)
P = network.init_solid_fermi_net_params(888, atoms=sc.original_cell.atom_coords(), spins=sc.nelec)
ld = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8, method_name="eval_logdet")
hp = ld.apply.hotpath()
X = torch.as_tensor(C.init_walkers(sc, B, seed=5)).cuda()
el = hamiltonian.local_energy_seperate(ld.apply, sc)
hp.debug_set("fused_digits", 0)
k0, _ = el(P, X); torch.cuda.synchronize()
print("unfused ok", complex(k0[0]), flush=True)
hp.debug_set("fused_digits", 1)
try:
    k1, _ = el(P, X); torch.cuda.synchronize()
    print("fused ok", complex(k1[0]), "max diff", float((k1 - k0).abs().max()), flush=True)
except Exception as e:
    print("FUSED FAILED:", type(e).__name__, str(e)[:500], flush=True)

"""Launch list of value-only forwards (the Metropolis inner loop: log|psi| of a proposal) at the benchmark configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepsolid_b200 import cell as C, network
sc = C.build_system("graphite54"); kl = C.make_klist(sc)
P = network.init_solid_fermi_net_params(888, atoms=sc.original_cell.atom_coords(), spins=sc.nelec)
net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8,
                                   method_name="eval_slogdet")
hp = net.apply.hotpath(); hp.set_params(P)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = torch.as_tensor(C.init_walkers(sc, B, seed=1)).cuda()
for _ in range(3):
    la, ph = hp.logpsi(X)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    la, ph = hp.logpsi(X)
e1.record(); torch.cuda.synchronize()
print(f"log psi of {B} walkers: {e0.elapsed_time(e1) / 5:.2f} ms")

"""deepsolid_b200: B200-native local-energy hot path of bytedance/DeepSolid.

Host side mirrors the reference's closures (``network.make_solid_fermi_net``,
``hamiltonian.local_energy_seperate``, ``qmc.make_mcmc_step``, ``train.make_loss``);
the compute path is hand-written CUDA for sm_100a behind the C ABI in
``include/deepsolid_b200.h`` (``libdeepsolid_b200.so``).
"""
__version__ = "0.1.0"

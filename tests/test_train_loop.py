"""End-to-end use of the closures the way the reference driver uses them (process.py:249-372): Metropolis sweep ->
energy gradient -> KFAC step, repeated.  The variational energy of a randomly initialised H4 chain must go down."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


@pytest.mark.gpu
def test_vmc_loop_lowers_the_energy(tmp_path):
    import train_vmc
    from deepsolid_b200 import checkpoint
    hist, _ = train_vmc.run(system="h4", batch=256, iterations=40, burn_in=10, log=None, ckpt_dir=str(tmp_path),
                            structure_factor=True)
    e = np.array([h["energy"] for h in hist])
    assert np.all(np.isfinite(e)) and all(np.isfinite(h["variance"]) for h in hist)
    assert all(0.0 < h["pmove"] <= 1.0 for h in hist)
    assert e[-8:].mean() < e[:8].mean() - 3.0 * e[:8].std() / np.sqrt(8), (e[:8], e[-8:])
    assert hist[-1]["structure_factor"].shape == (8,)
    last = checkpoint.find_last_checkpoint(str(tmp_path))
    assert last is not None and last.endswith("qmcjax_ckpt_000040.npz")

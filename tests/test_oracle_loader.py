"""CPU: the multi-sweep assembly oracle reproduces the reference `LoadPointCloudFromFile` (tests/golden/loader.npz,
written by oracle/gen_golden.py from /root/reference) bit for bit."""
import os

import numpy as np

from oracle import loader_ref as LR


def test_loader_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "loader.npz"))
    for case in ("a", "b"):
        key, sweeps = LR.synth_sweeps(int(g["seed_" + case]))
        got = LR.assemble_ref(key, [sweeps[i] for i in g["order_" + case]])
        want = g["combined_" + case]
        assert got.dtype == want.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got, want)
        assert (np.abs(want[len(key):, 0]) >= 1.0).any() and len(want) < len(key) + sum(len(s[0]) for s in sweeps)

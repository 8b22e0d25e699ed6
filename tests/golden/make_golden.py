"""Generate the golden fixtures (run once, on CPU): python tests/golden/make_golden.py

The reference (TF1 / Python 2) cannot run here and ships no vectors of its own,
so these are outputs of the fp64 twin of oracle/l2hmc_oracle.py on fixed inputs.  They pin the oracle
against drift and give the GPU tests a target that does not depend on the oracle code at test time.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))

import golden_io  # noqa: E402
import util as U  # noqa: E402

FIXTURES = [
    # (file, config, n, seed, regime)
    ("c1_scg2_n200_init", "c1_scg2", 200, 11, "init"),      # BASELINE config 1 as in SCGExperiment.ipynb
    ("c1_scg2_n200_stress", "c1_scg2", 200, 12, "stress"),
    ("c2_scg50_n96_stress", "c2_scg50", 96, 13, "stress"),  # BASELINE config 2, reduced chain count
    ("c3_mog2_n128_stress", "c3_mog2", 128, 14, "stress"),
    ("c4_rw32_n96_stress", "c4_rw32", 96, 15, "stress"),
]

if __name__ == "__main__":
    for fname, cfg, n, seed, regime in FIXTURES:
        path = os.path.join(HERE, fname + ".npz")
        golden_io.save(path, fname, U.CONFIGS[cfg], n, seed, regime)
        print(path, os.path.getsize(path))
    hmc = dict(kind="gaussian", D=2, T=10, eps=0.15, hmc=True)  # notebook_utils.get_hmc_samples settings
    path = os.path.join(HERE, "hmc_scg2_n200.npz")
    golden_io.save(path, "hmc_scg2_n200", hmc, 200, 16, "init")
    print(path, os.path.getsize(path))
    # BASELINE config 5 in miniature: decoder-Bernoulli target + aux-conditioned nets (mnist_vae.py:104-178)
    path = os.path.join(HERE, "c5_vae_mini_n96.npz")
    golden_io.save_vae(path, "c5_vae_mini_n96", U.VAE_CONFIGS["c5_vae_mini"], 96, 17)
    print(path, os.path.getsize(path))

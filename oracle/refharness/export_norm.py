"""TEST/DATA INFRASTRUCTURE -- converts the reference's committed normalisation statistics into the
small .npz inputs shipped under ao_marl_b200/data/normalization/.

The statistics (per-mode mean/std of Btt commands and residuals, and the action bounds zn_norm) were
produced by the reference authors WITH the real COMPASS simulator
(src/reinforcement_learning/helper_functions/preprocessing/normalization/obtain_normalization.py:139-250);
they are inputs of the environment (ao_env.py:251-306, rlSupervisor.py:255-282), not code.
    python -m oracle.refharness.export_norm
"""
import os
import pickle

import numpy as np

REF = "/root/reference/src/reinforcement_learning/helper_functions/preprocessing/normalization/"
OUT = os.path.join(os.path.dirname(__file__), "..", "..", "ao_marl_b200", "data", "normalization")
NAMES = ["production_sh_10x10_2m", "production_sh_40x40_8m_3layers", "production_sh_40x40_8m_3layers_d1_noise",
         "production_sh_40x40_8m_3layers_same_dir_roket", "production_sh_40x40_8m_3layers_noise_M9"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for n in NAMES:
        with open(REF + "state_normalization/normalization_%s_zernike_space.pickle" % n, "rb") as f:
            p = pickle.load(f)
        out = {}
        for key in ("dm", "wfs", "dm_residual"):
            for stat in ("mean", "std", "max", "min"):
                out["%s_%s" % (key, stat)] = np.asarray(p[key][stat], dtype=np.float32)
        out["zn_norm"] = np.load(REF + "normalization_action_zernike/zn_norm_%s.npy" % n).astype(np.float32)
        np.savez_compressed(os.path.join(OUT, n + ".npz"), **out)
        print("wrote", n)


if __name__ == "__main__":
    main()

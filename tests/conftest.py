import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden10():
    return np.load(os.path.join(GOLDEN, "ref_tables_10x10.npz"))


@pytest.fixture(scope="session")
def golden40():
    return np.load(os.path.join(GOLDEN, "ref_tables_40x40.npz"))


@pytest.fixture(scope="session")
def golden_rl():
    return np.load(os.path.join(GOLDEN, "ref_rl.npz"))


@pytest.fixture(scope="session")
def static10():
    from ao_marl_b200 import tables
    from ao_marl_b200.config import load_config_from_file
    t = tables.build_static(load_config_from_file("production_sh_10x10_2m.py"))
    tables.build_basis(t)
    return t


@pytest.fixture(scope="session")
def oracle_tab10(static10):
    tab = static10.as_oracle_dict()
    tab["wfs_index"] = static10.wfs_index
    return tab


@pytest.fixture(scope="session")
def oracle_imat10(oracle_tab10, static10):
    from oracle import loop
    return loop.measure_imat(oracle_tab10, static10.p_pzt.push4imat, static10.p_tt.push4imat)

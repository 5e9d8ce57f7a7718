import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


T2_CASES = [
    "t2_seed42_P10_K5_u100",
    "t2_seed7_P100_K5_u60_o5",
    "t2_seed3_P1000_K20_u24_organic_count",
    "t2_seed5_P60_K8_u40_bandit_count_flips",
    "t2_seed9_P33_K3_u40_omega_bandits",
    "t2_seed13_P40_K6_u60_o3_organic_user_count",
    "t2_seed21_P500_K10_u30_organic_user_count",
]


# more reference goldens for the CPU oracle only (degenerate shapes, unusual transition matrices)
T2_CASES_ORACLE_ONLY = [
    "t2_seed31_P1_K2_u80",
    "t2_seed32_P7_K1_u80_o6",
    "t2_seed33_P20_K3_u120_fastleave",
    "t2_seed41_P25_K4_u60_eps_organic_count",   # the reference's EpsilonGreedy(OrganicCount), epsilon = 0.25
    "t2_seed51_P30_K4_u50_o4_normal_time",      # the reference's NormalTimeGenerator clock (mu 0.5, sigma 2)
]


def oracle_env_from_golden(g):
    from oracle import oracle as O
    return O.OracleEnv(g["Gamma"], g["mu_organic"], g["beta"], g["mu_bandit"], g["state_transition"],
                       float(g["sigma_omega_initial"]), float(g["sigma_omega"]), int(g["seed"]),
                       bool(int(g["change_omega_for_bandits"])),
                       normal_time=tuple(g["normal_time"]) if "normal_time" in g.files else None)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O

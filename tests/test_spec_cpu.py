"""CPU: the product's RNG/math spec (host instantiation of recogym_b200/csrc/rg_math.cuh exported
as rg_spec_*) equals the oracle bit-for-bit; the C ABI exports every declared symbol; host-side
parameter generation equals the literal reference (T1 goldens); error behaviour without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from recogym_b200 import _lib as L


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "recogym_b200.h")).read()
    declared = set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rg_env_desc", "rg_policy", "rg_rollout_out"}
    lib = L.lib()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export " + name
    assert set(L.EXPORTS) == declared
    assert lib.rg_version() >= 100


def test_spec_functions_equal_oracle_bitwise(oracle):
    lib = L.lib()
    rs = np.random.RandomState(123)
    for x in np.concatenate([rs.uniform(-720, 60, 4000), [0.0, -0.0, -700.0, -700.1, 1e-300]]):
        assert lib.rg_spec_exp(x) == oracle.exp(x)
    for x in np.exp(rs.uniform(-700, 700, 4000)):
        assert lib.rg_spec_log(x) == oracle.log(x)
    for p in np.concatenate([rs.uniform(0, 1, 4000), np.exp(rs.uniform(-700, -1, 2000)), [2.0 ** -53, 1 - 2.0 ** -53]]):
        assert lib.rg_spec_norm_ppf(p) == oracle.norm_ppf(p)
    for x in rs.uniform(-40, 40, 2000):
        assert lib.rg_spec_ff(x) == oracle.ff(x)
    out = np.zeros(4, np.uint32)
    for _ in range(500):
        seed, user = int(rs.randint(0, 2 ** 62)), int(rs.randint(0, 2 ** 40))
        t, slot = int(rs.randint(0, 2 ** 31)), int(rs.randint(0, 2 ** 10))
        lib.rg_spec_philox(seed, user, t, slot, out.ctypes.data)
        assert np.array_equal(out, oracle.philox(seed, user, t, slot))
        for w in (0, 1):
            assert lib.rg_spec_event_uniform(seed, user, t, w) == oracle.event_uniform(seed, user, t, w)
        K = int(rs.randint(1, 70))
        z = np.zeros(K)
        lib.rg_spec_normals(seed, user, t, K, z.ctypes.data)
        assert np.array_equal(z, oracle.normals(seed, user, t, K))
        P = int(rs.randint(1, 200000))
        assert lib.rg_spec_policy_uniform_action(seed, user, t, P) == oracle.policy_uniform_action(seed, user, t, P)


def test_philox_known_answer_on_product_lib():
    out = np.zeros(4, np.uint32)
    L.lib().rg_spec_philox(0xa4093822 | (0x299f31d0 << 32), 0x243f6a88 | (0x85a308d3 << 32), 0x13198a2e, 0x03707344,
                           out.ctypes.data)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("name,args", [
    ("t1_seed42_P10_K5", {"random_seed": 42}),
    ("t1_static_seed7_P100_K5_F0", {"random_seed": 7, "num_products": 100, "K": 5}),
    ("t1_static_seed11_P50_K8_F10", {"random_seed": 11, "num_products": 50, "K": 8, "number_of_flips": 10}),
])
def test_static_parameters_bit_identical_to_literal_reference(name, args):
    from recogym_b200 import Configuration, env_1_args
    from recogym_b200.envs.static_params import draw_static_params
    g = load_golden(name)
    sp = draw_static_params(Configuration({**env_1_args, **args}))
    assert np.array_equal(sp["Gamma"], g["Gamma"])
    assert np.array_equal(sp["mu_organic"].ravel(), g["mu_organic"])
    assert np.array_equal(sp["beta"], g["beta"])
    assert np.array_equal(sp["mu_bandit"].ravel(), g["mu_bandit"])
    assert np.array_equal(sp["state_transition"], g["state_transition"])
    if args.get("number_of_flips", 0) == 0:
        assert sp["beta"] is sp["Gamma"]  # alias, reco_env_v1.py:137


def test_no_gpu_means_loud_failure_not_fallback():
    import recogym_b200 as rg
    if L.device_count() > 0:
        pytest.skip("GPU present")
    env = rg.make("reco-gym-v1")
    with pytest.raises(rg.RecoGymB200Error) as ei:
        env.init_gym({**rg.env_1_args, "random_seed": 1})
    assert ei.value.code == L.RG_ENODEVICE


def test_argument_validation():
    lib = L.lib()
    d = L.EnvDesc()
    h = C.c_void_p()
    assert lib.rg_env_create(None, 0, C.byref(h)) == L.RG_EINVAL
    G = np.zeros((4, 2)); mu = np.zeros(4)
    d.num_products, d.K = 4, 2
    d.Gamma = d.beta = G.ctypes.data
    d.mu_organic = d.mu_bandit = mu.ctypes.data
    for i, v in enumerate([0.5, 0.4, 0.2, 0, 1, 0, 0, 0, 1]):   # first row sums to 1.1
        d.state_transition[i] = v
    assert lib.rg_env_create(C.byref(d), 0, C.byref(h)) == L.RG_EINVAL
    assert b"sum to 1" in lib.rg_last_error()
    d.state_transition[2] = 0.1
    d.K = 1000
    assert lib.rg_env_create(C.byref(d), 0, C.byref(h)) == L.RG_EINVAL
    d.K = 2
    d.num_products = 0
    assert lib.rg_env_create(C.byref(d), 0, C.byref(h)) == L.RG_EINVAL
    assert lib.rg_rollout_host(None, None, 0, 0, 0, None, 0) == L.RG_EINVAL


def test_value_types_follow_reference_contract():
    import copy
    import pickle
    import recogym_b200 as rg
    c = rg.Configuration({"a": 1, "b": [1, 2]})
    c.a = 5
    assert c.a == 1                                   # immutable (configuration.py:20-21)
    assert copy.deepcopy(c) is c                      # configuration.py:23-24
    assert pickle.loads(pickle.dumps(c)).b == [1, 2]
    s = rg.OrganicSessions()
    s.next(rg.DefaultContext(3, 7), 4)
    assert s == [{"t": 3, "u": 7, "z": "pageview", "v": 4}] and s.get_type() == "organic"
    o = rg.Observation(rg.DefaultContext(1, 2), s)
    assert o.context().time() == 1 and o.context().user() == 2 and o.sessions() is s and o.current_sessions is s
    a = rg.agents.Agent(c).act(o, 0, False)
    assert a == {"t": 1, "u": 2}
    assert set(rg.env_1_args) >= {"num_products", "K", "sigma_omega", "sigma_omega_initial", "number_of_flips",
                                  "sigma_mu_organic", "change_omega_for_bandits", "prob_leave_organic", "random_seed"}
    with pytest.raises(KeyError):
        rg.make("reco-gym-v7")


def test_count_agents_compile_to_tables():
    import recogym_b200 as rg
    from recogym_b200.policy import device_policy, probe_last_view_table
    P = 6
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    s = rg.OrganicSessions()
    for v in (1, 1, 3):
        s.next(rg.DefaultContext(0, 0), v)
    oc.train(rg.Observation(rg.DefaultContext(0, 0), s), None, None, True)
    dense = np.zeros((P, P)); b = np.array([0, 2, 0, 1, 0, 0.0]); dense += np.outer(b, b)
    assert np.array_equal(oc.co_counts, dense)        # == B B^T of organic_count.py:76-82
    pol = device_policy(oc, P)
    assert pol["policy"] == "table" and np.array_equal(pol["table"], dense.argmax(1))
    assert np.array_equal(probe_last_view_table(oc, P), pol["table"])   # probing act() agrees
    bc = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    bc.last_product_viewed = 2
    bc.train(rg.Observation(rg.DefaultContext(0, 0), s), {"a": 4}, 1)
    assert bc.ctr[2, 4] == 2 / 3 and bc.last_product_viewed == 3
    assert device_policy(bc, P)["table"][2] == 4
    assert device_policy(None, P) == {"policy": "uniform"}


def test_time_step_spec_equals_oracle_clock(oracle):
    """rg_spec_time_step (NormalTimeGenerator.new_time's step, normal_time_generator.py:25-28, Philox slot 2) == the
    oracle's clock normal, bit for bit; the product library and the oracle share no code."""
    from recogym_b200 import _lib as L
    f = L.lib().rg_spec_time_step
    for seed, user, mu, sigma in ((51, 0, 0.5, 2.0), (7, 123456789012, 0.0, 1.0), (2 ** 40 + 3, 5, -1.5, 0.25)):
        for call in (0, 1, 2, 17, 4095):
            assert f(seed, user, call, mu, sigma) == abs(mu + sigma * oracle.time_normal(seed, user, call))


def test_blocked_flip_search_equals_the_reference_walk():
    """static_params.flip_index_blocked (row maxima instead of the P x P argsort: what number_of_flips > 0 needs at
    P = 100 000) == flip_index, the literal restatement of reco_env_v1.py:142-165 that is pinned to the reference,
    including when the flips exhaust the catalogue; and it runs at a size where the dense form cannot."""
    from recogym_b200.envs.static_params import flip_index, flip_index_blocked
    for P, K, flips, seed in ((50, 8, 10, 11), (333, 5, 40, 2), (1200, 20, 150, 3), (64, 4, 40, 4), (7, 3, 3, 5)):
        G = np.random.RandomState(seed).normal(size=(P, K))
        a, b = flip_index(G, flips), flip_index_blocked(G, flips, block=97)
        assert np.array_equal(a, b), (P, flips)
        assert (a != np.arange(P)).sum() == 2 * min(flips, P // 2)
    G = np.random.RandomState(9).normal(size=(20000, 16))
    idx = flip_index_blocked(G, 25)
    moved = np.flatnonzero(idx != np.arange(20000))
    assert len(moved) == 50 and np.array_equal(idx[idx[moved]], moved)

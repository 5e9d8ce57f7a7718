"""reco-gym-v0 on the GPU, through the C ABI (rg_env0_*): the CUDA rollout against the reference's own logs (T2 goldens,
tests/golden/make_golden.py --v0-only) and against the oracle at sizes the oracle finishes in seconds; the gym facade
(`make('reco-gym-v0')`: generate_logs, reset / step / step_offline by replay) against the same goldens."""
import numpy as np
import pytest

from conftest import load_golden
from test_env0_cpu import PARAMS, V0_CASES, golden_params, golden_rows

pytestmark = pytest.mark.gpu


def device_env0(g):
    from recogym_b200.envs.reco_env_v0 import DeviceEnv0, device_tables
    return DeviceEnv0(device_tables(golden_params(g)), int(g["seed"]))


@pytest.mark.parametrize("case", V0_CASES)
def test_v0_rollout_equals_reference_logs_and_oracle(case):
    from oracle import v0_oracle as V
    g = load_golden(case)
    sp, seed, P = golden_params(g), int(g["seed"]), int(g["product_transition"].shape[0])
    de = device_env0(g)
    table = g["table"] if "table" in g.files else None
    r = de.rollout(int(g["n_users"]), str(g["policy"]), table, 0, int(g["n_organic_users"]))
    want = golden_rows(g)
    for k in want:
        assert np.array_equal(r[k], want[k]), (case, k)
    rs = np.random.RandomState(4)
    tb = rs.randint(0, P, P).astype(np.int32)
    for pol, t, ub, n, no in (("uniform", None, 1000, 1500, 0), ("table", tb, 123456789012, 900, 17), ("uniform", None, 5, 0, 9)):
        r = de.rollout(n, pol, t, ub, no, rows_per_user=2.0)                       # forces the capacity retry
        o = V.rollout(sp, seed, ub, n, pol, t, no)
        for k in ("u", "e", "z", "x", "c", "phantom", "n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "state", "totals"):
            assert np.array_equal(r[k], o[k]), (case, pol, k)
    acts = g["trace_actions"]
    for k in (0, 1, len(acts) // 2, len(acts)):
        r = de.rollout(1, "replay", acts[:k], int(g["trace_user"]), 0)
        o = V.rollout(sp, seed, int(g["trace_user"]), 1, "replay", None, 0, actions=acts[:k])
        for key in ("e", "z", "x", "c", "state", "last_view"):
            assert np.array_equal(r[key], o[key]), (case, k, key)
    de.close()


def test_v0_counters_at_scale_are_shard_invariant_and_plausible():
    """1M users in one launch: totals == the sum over four shards (a user is a pure function of its id), events per user
    close to the geometric law of the transition matrix, re-seeding changes the draws."""
    g = load_golden(V0_CASES[0])
    de = device_env0(g)
    whole = de.rollout(1_000_000, "uniform", None, 0, 0, log=False, per_user=False)["totals"]
    parts = sum(de.rollout(250_000, "uniform", None, 250_000 * k, 0, log=False, per_user=False)["totals"] for k in range(4))
    assert np.array_equal(whole, parts) and whole[0] == 1_000_000 and whole[5] == 1_000_000
    assert 95.0 < whole[1] / whole[0] < 106.0          # 1 / prob_leave = 100 events per user
    assert 0.0 < whole[4] / whole[3] < 0.2
    first = de.rollout(250_000, "uniform", None, 0, 0, log=False, per_user=False)["totals"]
    de.set_seed(int(g["seed"]) + 1)                     # reset_random_seed(epoch = 1)
    assert not np.array_equal(de.rollout(250_000, "uniform", None, 0, 0, log=False, per_user=False)["totals"], first)
    de.close()


class _TableAgent:
    """an agent that compiles to a last-view table (what the count agents' device_policy() returns)"""

    def __init__(self, table, ps_table):
        self.table, self.ps_table = table, ps_table

    def device_policy(self):
        return {"policy": "table", "table": self.table, "ps_table": self.ps_table}

    def reset(self):
        pass


@pytest.mark.parametrize("case", V0_CASES)
def test_v0_facade_generate_logs_and_gym_protocol_equal_the_reference(case):
    import recogym_b200 as rg
    g = load_golden(case)
    P = int(g["product_transition"].shape[0])
    over = {"t2_v0_seed42_P10_u120_o5": {}, "t2_v0_seed7_P30_c3_u150": {"phi_var": 4.0, "prob_leave_organic": 0.03},
            "t2_v0_seed9_P12_c4_bandit_count": {"phi_var": 40.0}}[case]
    env = rg.make("reco-gym-v0")
    env.init_gym({**rg.env_0_args, "random_seed": int(g["seed"]), "num_products": P, "num_clusters": int(g["num_clusters"]), **over})
    for k in PARAMS + ("phi",):
        assert np.array_equal(getattr(env, k), g[k]), k
    agent = _TableAgent(g["table"], g["ps_table"]) if "table" in g.files else None
    df = env.generate_logs(int(g["n_users"]), agent=agent, num_organic_offline_users=int(g["n_organic_users"]))
    assert list(df.columns) == ["t", "u", "z", "v", "a", "c", "ps", "ps-a"] and len(df) == len(g["log_u"])
    z = (df["z"].to_numpy() == "bandit").astype(np.int8)
    col = lambda name: np.nan_to_num(df[name].to_numpy(dtype="float64", na_value=np.nan), nan=-1)
    assert np.array_equal(z, g["log_z"]) and (df["t"] == 0).all()
    for name in ("u", "v", "a", "c"):
        assert np.array_equal(col(name), g["log_" + name].astype(np.float64)), name
    assert np.array_equal(np.nan_to_num(df["ps"].to_numpy(), nan=-1), np.nan_to_num(g["log_ps"], nan=-1))
    # gym protocol under GIVEN actions: the observations' views, rewards and done flags of the reference's step()
    env.reset(int(g["trace_user"]))
    obs, reward, done, _ = env.step(None)
    views, rewards, dones = [int(s["v"]) for s in obs.sessions()], [], [bool(done)]
    assert reward is None
    for a in g["trace_actions"]:
        assert not done
        obs, reward, done, _ = env.step(int(a))
        views += [int(s["v"]) for s in obs.sessions()]
        rewards.append(int(reward)); dones.append(bool(done))
    assert np.array_equal(views, g["trace_views"]) and np.array_equal(rewards, g["trace_rewards"])
    assert np.array_equal(np.array(dones, np.int8), g["trace_done"])
    # step_offline loop with the default uniform policy == the fused generate_logs rows of the same users
    if agent is None:
        b = env._generate_logs_stepwise(6, None, 2)
        a = env.generate_logs(6, num_organic_offline_users=2)
        for name in ("u", "v", "a", "c", "ps"):
            xa = np.nan_to_num(a[name].to_numpy(dtype="float64", na_value=np.nan), nan=-1)
            xb = np.nan_to_num(b[name].to_numpy(dtype="float64", na_value=np.nan), nan=-1)
            assert np.array_equal(xa, xb), name
        assert (a["z"] == b["z"]).all()

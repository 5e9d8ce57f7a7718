"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the
committed reference fixtures (T2 = the reference's own code driven by the same Philox streams).

Bar: every discrete output (per-user event/organic/bandit/click counts, last view, event hash,
histograms, session contents, rewards, states) is bit-exact; omega is bit-exact as well because
the device and the oracle evaluate the same binary64 operation sequence.
"""
import numpy as np
import pytest

from conftest import T2_CASES, load_golden, oracle_env_from_golden

pytestmark = pytest.mark.gpu

MODES = {"exact": 0, "fast": 1, "tensor": 2}


def device_env_from_golden(g, device=0):
    from recogym_b200 import DeviceEnv
    sp = {k: g[k] for k in ("Gamma", "mu_organic", "beta", "mu_bandit", "state_transition")}
    return DeviceEnv(sp, float(g["sigma_omega_initial"]), float(g["sigma_omega"]), int(g["seed"]),
                     bool(int(g["change_omega_for_bandits"])), device)


def synthetic_env(P, K, seed, flips=0, **over):
    """host-side static parameters exactly as RecoEnv1.init_gym would draw them."""
    from recogym_b200 import Configuration, env_1_args
    from recogym_b200.envs.static_params import draw_static_params
    cfg = Configuration({**env_1_args, "num_products": P, "K": K, "random_seed": seed, "number_of_flips": flips,
                         **over})
    return cfg, draw_static_params(cfg)


def both_envs(P, K, seed, **over):
    from oracle import oracle as O
    from recogym_b200 import DeviceEnv
    cfg, sp = synthetic_env(P, K, seed, **over)
    oe = O.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                     cfg.sigma_omega_initial, cfg.sigma_omega, seed, cfg.change_omega_for_bandits)
    de = DeviceEnv(sp, cfg.sigma_omega_initial, cfg.sigma_omega, seed, cfg.change_omega_for_bandits)
    return oe, de


PER_USER = ("n_events", "n_organic", "n_bandit", "n_clicks", "last_view")


def assert_rollout_equal(r_gpu, r_cpu, hist=True, omega=True, ambiguous_ok=False):
    """totals[7] counts the draws that were re-drawn in the oracle's own operation order (tier 4): correct by
    construction, but at these sizes none is expected unless the test forces the path (ambiguous_ok)."""
    for k in PER_USER:
        assert np.array_equal(r_gpu[k], r_cpu[k]), k
    assert np.array_equal(np.asarray(r_gpu["hash"]).view(np.uint64), r_cpu["hash"]), "hash"
    assert np.array_equal(r_gpu["totals"][:6], r_cpu["totals"][:6]), (r_gpu["totals"], r_cpu["totals"])
    assert ambiguous_ok or r_gpu["totals"][7] == 0, "oracle-order re-draws at a size where none is expected"
    if hist:
        for k in ("view_hist", "action_hist", "click_hist"):
            assert np.array_equal(r_gpu[k], r_cpu[k]), k
    if omega:
        assert np.array_equal(r_gpu["omega"], r_cpu["omega"]), "omega"


def test_device_math_equals_host_and_oracle(oracle):
    import ctypes as C
    import torch
    from recogym_b200 import _lib as L
    rs = np.random.RandomState(5)
    cases = {
        0: np.concatenate([rs.uniform(-720, 60, 20000), [0.0, -700.0, -700.5]]),
        1: np.exp(rs.uniform(-700, 700, 20000)),
        2: np.concatenate([rs.uniform(0, 1, 20000), np.exp(rs.uniform(-700, -1, 5000)), [2.0 ** -53, 1 - 2.0 ** -53]]),
        3: rs.uniform(-40, 40, 20000),
    }
    fn = {0: oracle.exp, 1: oracle.log, 2: oracle.norm_ppf, 3: oracle.ff}
    for kind, x in cases.items():
        xi = torch.as_tensor(x, device="cuda")
        yo = torch.empty_like(xi)
        L.check(L.lib().rg_spec_eval_device(kind, xi.numel(), C.c_void_p(xi.data_ptr()), C.c_void_p(yo.data_ptr()), None))
        y = yo.cpu().numpy()
        ref = np.array([fn[kind](v) for v in x])
        assert np.array_equal(y, ref), kind
    n = 4000
    trip = np.stack([rs.randint(0, 2 ** 50, n).astype(np.float64), rs.randint(0, 2 ** 40, n).astype(np.float64),
                     rs.randint(0, 2 ** 31, n).astype(np.float64)], 1)
    ti = torch.as_tensor(trip.reshape(-1), device="cuda")
    for kind in (4, 5, 6):
        yo = torch.empty(2 * n, dtype=torch.float64, device="cuda")
        L.check(L.lib().rg_spec_eval_device(kind, n, C.c_void_p(ti.data_ptr()), C.c_void_p(yo.data_ptr()), None))
        y = yo.cpu().numpy().reshape(n, 2)
        for i in range(0, n, 7):
            s, u, t = (int(v) for v in trip[i])
            if kind == 4:
                z = oracle.normals(s, u, t, 2)
                assert y[i, 0] == z[0] and y[i, 1] == z[1]
            elif kind == 5:
                assert y[i, 0] == oracle.event_uniform(s, u, t, 0) and y[i, 1] == oracle.event_uniform(s, u, t, 1)
            else:
                assert y[i, 0] == oracle.policy_uniform_action(s, u, t, 1000003)
                assert y[i, 1] == float(oracle.philox(s, u, t, 1)[3])


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
@pytest.mark.parametrize("case", T2_CASES)
def test_rollout_equals_oracle_and_reference_logs(oracle, case, mode):
    """Golden cases: CUDA rollout == C oracle per user, and its counters == the T2 reference logs."""
    g = load_golden(case)
    oe = oracle_env_from_golden(g)
    de = device_env_from_golden(g)
    table = g["table"] if "table" in g.files else None
    n, no = int(g["n_users"]), int(g["n_organic_users"])
    rc = oracle.rollout(oe, 0, n, policy=str(g["policy"]), table=table, n_organic_users=no, hist=True, omega=True)
    rg = de.rollout_host(n, str(g["policy"]), table, 0, no, MODES[mode], per_user=True, omega=True, hist=True)
    assert_rollout_equal(rg, rc)
    # directly against the reference's own log columns
    u, z, c, v = g["log_u"], g["log_z"], g["log_c"], g["log_v"]
    assert np.array_equal(rg["n_organic"], np.bincount(u[z == 0], minlength=n + no))
    assert np.array_equal(rg["n_bandit"] + (np.arange(n + no) >= no), np.bincount(u[z == 1], minlength=n + no))
    assert np.array_equal(rg["n_clicks"], np.bincount(u[c == 1], minlength=n + no))
    assert np.array_equal(rg["view_hist"], np.bincount(v[z == 0], minlength=oe.P))
    de.close()


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
@pytest.mark.parametrize("P,K,n,policy", [
    (1, 1, 300, "uniform"),        # single product: view always 0
    (2, 3, 400, "table"),
    (37, 7, 500, "uniform"),       # P not a multiple of 32, odd K
    (1025, 20, 200, "table"),      # one product past a super-chunk boundary
    (3000, 33, 64, "uniform"),     # K > 32
    (10000, 20, 48, "table"),      # BASELINE config C3 shape
    (2048, 64, 40, "uniform"),     # K = 64 (C4 latent size)
])
def test_rollout_shapes_and_edges(oracle, P, K, n, policy, mode):
    oe, de = both_envs(P, K, 1000 + P + K)
    table = (np.random.RandomState(P).randint(0, P, P).astype(np.int32)) if policy == "table" else None
    rc = oracle.rollout(oe, 5, n, policy=policy, table=table, n_organic_users=3, hist=True, omega=True, threads=8)
    rg = de.rollout_host(n, policy, table, 5, 3, MODES[mode], per_user=True, omega=True, hist=True)
    assert_rollout_equal(rg, rc)
    de.close()


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_rollout_model_variants(oracle, mode):
    for over in ({"change_omega_for_bandits": True, "sigma_omega": 0.4},
                 {"flips": 15, "normalize_beta": True},
                 {"prob_leave_organic": 0.2, "prob_organic_to_bandit": 0.5, "prob_bandit_to_organic": 0.6},
                 {"sigma_omega_initial": 0.0, "sigma_omega": 0.0}):
        oe, de = both_envs(64, 6, 77, **over)
        rc = oracle.rollout(oe, 0, 300, hist=True, omega=True, threads=4)
        rg = de.rollout_host(300, "uniform", None, 0, 0, MODES[mode], per_user=True, omega=True, hist=True)
        assert_rollout_equal(rg, rc)
        de.close()


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_user_views_policy_equals_oracle(oracle, mode):
    """RG_POLICY_USER_VIEWS (OrganicUserEventCounterAgent: action ~ the user's own view counts) against the
    oracle's dense restatement: small P (every product viewed, many repeats), P past a chunk boundary, the C3
    shape, and change_omega_for_bandits (the serial bandit loop of the CTA kernel)."""
    for P, K, n, over in ((7, 3, 400, {}), (1025, 20, 150, {}), (10000, 20, 40, {}),
                          (64, 6, 300, {"change_omega_for_bandits": True, "sigma_omega": 0.4}),
                          (50, 4, 300, {"prob_organic_to_bandit": 0.1, "prob_leave_organic": 0.002,
                                        "prob_leave_bandit": 0.002})):   # long organic sessions: long lists
        oe, de = both_envs(P, K, 300 + P, **over)
        rc = oracle.rollout(oe, 2, n, policy="user_views", n_organic_users=2, hist=True, omega=True, threads=8)
        rg = de.rollout_host(n, "user_views", None, 2, 2, MODES[mode], per_user=True, omega=True, hist=True)
        assert_rollout_equal(rg, rc)
        assert rg["totals"][9] == 0          # no view-list overflow
        assert (rg["action_hist"] > 0).sum() <= (rg["view_hist"] > 0).sum()   # only viewed products are recommended
        de.close()


def test_empty_and_device_entry_points(oracle):
    import torch
    oe, de = both_envs(50, 5, 9)
    r0 = de.rollout_host(0, "uniform", None, 0, 0, 0)
    assert r0["totals"].sum() == 0 and r0["n_events"].shape == (0,)
    # device-resident call == host call; accumulation of totals across calls; seed re-keying
    r1 = de.rollout_device(200, "uniform", None, 0, 0, 0, per_user=True, omega=True, hist=True)
    r2 = de.rollout_host(200, "uniform", None, 0, 0, 0, per_user=True, omega=True, hist=True)
    torch.cuda.synchronize()
    for k in PER_USER + ("hash", "omega", "view_hist", "totals"):
        assert np.array_equal(r1[k].cpu().numpy(), np.asarray(r2[k]).view(r1[k].cpu().numpy().dtype)), k
    tot = r1["totals"]
    de.rollout_device(100, "uniform", None, 200, 0, 0, per_user=False, totals=tot)
    torch.cuda.synchronize()
    r3 = de.rollout_host(300, "uniform", None, 0, 0, 0)
    assert np.array_equal(tot.cpu().numpy(), r3["totals"])
    de.set_seed(10)                                   # reset_random_seed(epoch=1)
    oe.c.seed = 10
    rc = oracle.rollout(oe, 0, 100, hist=True, omega=True)
    assert_rollout_equal(de.rollout_host(100, "uniform", None, 0, 0, 0, per_user=True, omega=True, hist=True), rc)
    de.close()


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_sharding_invariance_and_full_size_properties(mode):
    """BASELINE config C2 size (P=100, K=5, 1M users): results do not depend on how users are split
    over launches/ranks; counters are self-consistent; events/user follows the geometric law."""
    from recogym_b200 import DeviceEnv
    cfg, sp = synthetic_env(100, 5, 42)
    de = DeviceEnv(sp, 1, 0.1, 42)
    N = 1_000_000 if mode != "exact" else 100_000
    full = de.rollout_host(N, "uniform", None, 0, 0, MODES[mode], per_user=True, hist=True)
    parts = [de.rollout_host(N // 4, "uniform", None, r * (N // 4), 0, MODES[mode], per_user=True, hist=True)
             for r in range(4)]
    for k in PER_USER + ("hash",):
        assert np.array_equal(full[k], np.concatenate([p[k] for p in parts])), k
    for k in ("view_hist", "action_hist", "click_hist", "totals"):
        assert np.array_equal(full[k][:10], sum(p[k] for p in parts)[:10]), k     # totals[10:] are cycle diagnostics
    t = full["totals"]
    assert t[0] == N and t[5] == N
    assert t[1] == full["n_events"].sum() == t[2] + t[3]
    assert t[2] == full["view_hist"].sum() and t[3] == full["action_hist"].sum()
    assert t[4] == full["click_hist"].sum() == full["n_clicks"].sum()
    assert t[7] == 0
    # every event leaves with prob .01 unless a click overrides -> mean a bit above 100
    assert 99.0 < full["n_events"].mean() < 104.0
    ctr = t[4] / t[3]
    assert 0.005 < ctr < 0.03
    de.close()


def test_c3_shape_fast_equals_exact_on_device():
    """P=10000, K=20 (BASELINE C3) at a size the CPU oracle cannot do in seconds: the fast sampler
    (fp32 + ambiguity filter + binary64 fallback) must equal the all-binary64 sampler bit-for-bit."""
    from recogym_b200 import DeviceEnv
    cfg, sp = synthetic_env(10000, 20, 42)
    de = DeviceEnv(sp, 1, 0.1, 42)
    table = np.random.RandomState(1).randint(0, 10000, 10000).astype(np.int32)
    a = de.rollout_host(3000, "table", table, 0, 0, 0, per_user=True, hist=True, omega=True)
    for mode in (1, 2):
        b = de.rollout_host(3000, "table", table, 0, 0, mode, per_user=True, hist=True, omega=True)
        for k in PER_USER + ("hash", "view_hist", "action_hist", "click_hist", "omega"):
            assert np.array_equal(a[k], b[k]), (mode, k)
        assert np.array_equal(a["totals"][:6], b["totals"][:6])
        assert a["totals"][7] == 0 and b["totals"][7] == 0
        assert 0 < b["totals"][8] <= b["totals"][6] < 0.2 * b["totals"][2]   # tier 3 << tier 2 << organic draws
    de.close()


def test_c4_shape_fast_equals_exact_on_device():
    """P=100000, K=64 (BASELINE C4): RG_MATH_FAST needs its global-memory spill of the re-sampling sums and of the omega
    rows (one CTA per SM); RG_MATH_TENSOR runs the contraction on tcgen05 in three K slabs per product tile.  Discrete
    results must equal the all-binary64 sampler (the oracle comparison at this shape: tests/test_gpu_bench_scale.py)."""
    from recogym_b200 import DeviceEnv
    cfg, sp = synthetic_env(100000, 64, 42)
    de = DeviceEnv(sp, 1, 0.1, 42)
    table = np.random.RandomState(1).randint(0, 100000, 100000).astype(np.int32)
    a = de.rollout_host(400, "table", table, 0, 0, 0, per_user=True, hist=True, omega=True)
    for mode in (1, 2):
        b = de.rollout_host(400, "table", table, 0, 0, mode, per_user=True, hist=True, omega=True)
        for k in PER_USER + ("hash", "view_hist", "action_hist", "click_hist", "omega"):
            assert np.array_equal(a[k], b[k]), (mode, k)
        assert np.array_equal(a["totals"][:6], b["totals"][:6])
    de.close()


def test_batch_protocol_equals_oracle(oracle):
    """rg_batch_reset/step/continue == the oracle's reset/step, sessions included (tiny sess_cap
    forces the continuation path)."""
    from recogym_b200 import BatchState
    oe, de = both_envs(40, 5, 21, prob_leave_organic=0.06)   # short episodes: the host loop is per step
    N = 300
    uids = np.arange(3, 3 + N)
    st, off, sv = oracle.batch_reset(oe, uids)
    bs = BatchState(de, uids, sess_cap=3)

    def gather(first):
        """run continuation launches until no user is left organic; returns per-user view lists"""
        views = [[] for _ in range(N)]
        h = bs.read()
        while True:
            for i in range(N):
                views[i].extend(h["sess_v"][i, :h["sess_len"][i]].tolist())
            if not (h["state"] == 0).any():
                return views, h
            assert int(bs.n_unfinished.item()) == int((h["state"] == 0).sum())
            bs.continue_sessions()
            h = bs.read()

    bs.reset()
    views, h = gather(True)
    for it in range(3000):
        for i in range(N):
            assert views[i] == sv[off[i]:off[i + 1]].tolist(), (it, i)
        assert np.array_equal(h["state"], st["state"]) and np.array_equal(h["t"], st["t"])
        assert np.array_equal(h["last_view"], st["last_view"])
        assert np.array_equal(bs.omega.cpu().numpy(), st["omega"])
        if (st["state"] == 2).all():
            break
        act = np.array([oracle.policy_uniform_action(21, int(u), int(t), oe.P) for u, t in zip(uids, st["t"])], np.int32)
        rew, off, sv = oracle.batch_step(oe, st, act)
        bs.step(act)
        views, h = gather(False)
        assert np.array_equal(h["reward"], rew), it
    assert (st["state"] == 2).all()
    de.close()


def test_facade_reproduces_reference_logs_and_trace():
    """The reference-facing object (make('reco-gym-v1') ... generate_logs) yields the T2 reference
    DataFrame columns; the 'Getting Started' step_offline loop works unchanged."""
    import recogym_b200 as rg
    g = load_golden("t2_seed42_P10_K5_u100")
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 42})
    assert np.array_equal(env.Gamma, g["Gamma"]) and np.array_equal(env.state_transition, g["state_transition"])
    df = env.generate_logs(25)
    m = g["log_u"] < 25
    assert list(df.columns) == ["t", "u", "z", "v", "a", "c", "ps", "ps-a"]
    assert np.array_equal(df["u"].to_numpy(dtype=np.int64), g["log_u"][m])
    assert np.array_equal(df["t"].to_numpy().astype(np.int64), g["log_t"][m])
    assert np.array_equal((df["z"] == "bandit").to_numpy().astype(np.int8), g["log_z"][m])
    for col, key in (("v", "log_v"), ("a", "log_a"), ("c", "log_c")):
        x = df[col].to_numpy(dtype="float64", na_value=np.nan)
        assert np.array_equal(np.where(np.isnan(x), -1, x).astype(np.int64), g[key][m].astype(np.int64)), col
    # gym protocol by hand (Getting Started.ipynb cell 3 shape)
    env.reset(7)
    obs, reward, done, info = env.step(None)
    assert reward is None and info == {} and len(obs.sessions()) >= 1 and obs.sessions()[0]["t"] == 0
    steps = 0
    while not done:
        action, obs, reward, done, info = env.step_offline(obs, reward, done)
        assert reward in (0, 1) and 0 <= action["a"] < 10 and action["ps"] == 0.1
        steps += 1
    assert env.state == 2 and steps >= 0
    # a table agent through the facade == the fused rollout
    g2 = load_golden("t2_seed5_P60_K8_u40_bandit_count_flips")
    env2 = rg.make("reco-gym-v1")
    env2.init_gym({**rg.env_1_args, "random_seed": 5, "num_products": 60, "K": 8, "number_of_flips": 12,
                   "sigma_omega": 0.3})
    assert np.array_equal(env2.beta, g2["beta"])
    agent = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": 60}))
    agent.ctr[np.arange(60), g2["table"]] = 0.9          # make argmax == the golden table
    df2 = env2.generate_logs(40, agent=agent)
    a = df2["a"].to_numpy(dtype="float64", na_value=np.nan)
    assert np.array_equal(np.where(np.isnan(a), -1, a).astype(np.int64), g2["log_a"].astype(np.int64))
    cnt = env2.generate_logs_counters(40, "table", g2["table"], math_mode=1)
    assert cnt["successes"] == int((g2["log_c"] == 1).sum())
    assert cnt["failures"] == int((g2["log_c"] == 0).sum())


def test_drivers_test_agent_and_verify_agents():
    import copy
    import recogym_b200 as rg
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 42, "num_products": 20})
    e2 = copy.deepcopy(env)
    assert e2.device_env is not env.device_env and np.array_equal(e2.Gamma, env.Gamma)
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": 20}))
    q = rg.test_agent(copy.deepcopy(env), copy.deepcopy(oc), 30, 2000, 10)
    q_again = rg.test_agent(copy.deepcopy(env), copy.deepcopy(oc), 30, 2000, 10)
    assert q == q_again and 0.0 < q[1] < q[0] < q[2] < 0.1          # deterministic_test.py:86-94 property
    ra = rg.agents.RandomAgent(rg.Configuration({**rg.agents.random_args, "num_products": 20, "random_seed": 1}))
    df = rg.verify_agents(env, 3000, {"random": ra, "organic": oc})
    assert list(df.columns) == ["Agent", "0.025", "0.500", "0.975"] and len(df) == 2
    assert ((df["0.025"] < df["0.500"]) & (df["0.500"] < df["0.975"])).all()


def test_tensor_core_logits_match_binary64():
    """tcgen05 contraction (fp16 hi/lo operand split, three UMMA terms, fp32 accumulation) against binary64 logits: validates the
    UMMA descriptors / operand layout / TMEM read-back and measures the error the tensor-core sampler's
    filter has to cover."""
    import ctypes as C
    import torch
    from recogym_b200 import DeviceEnv, _lib as L
    for P, K in ((1000, 20), (10000, 20), (300, 5), (700, 64)):
        cfg, sp = synthetic_env(P, K, 7 + P)
        de = DeviceEnv(sp, 1, 0.1, 1)
        rs = np.random.RandomState(P)
        omega = rs.normal(size=(128, K))
        omega[0] = 0.0
        omega[1] *= 3.0
        G, mu = sp["Gamma"], sp["mu_organic"].ravel()
        log2e = 1.4426950408889634
        logits = (omega @ G.T + mu[None, :]) * log2e
        shift = (np.linalg.norm(omega, axis=1) * np.linalg.norm(G, axis=1).max() + mu.max()) * log2e - 100.0
        shift = shift.astype(np.float32)
        nsub = (P + 127) // 128
        om_d = torch.as_tensor(omega, device="cuda")
        sh_d = torch.as_tensor(shift, device="cuda")
        out = torch.zeros((128, nsub * 128), dtype=torch.float32, device="cuda")
        L.check(L.lib().rg_tc_logits_device(de._h, C.c_void_p(om_d.data_ptr()), C.c_void_p(sh_d.data_ptr()),
                                            C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.float64)
        ref = logits - shift.astype(np.float64)[:, None]
        err = np.abs(got[:, :P] - ref)
        scale = np.abs(ref).max()
        print("tc logits P=%d K=%d: max abs err %.3e (log2 units), max |ref| %.1f" % (P, K, err.max(), scale))
        assert err.max() < 2e-4, (P, K, err.max())
        assert (got[:, P:] < -5e4).all()           # padded products carry no mass (ex2 of their logit is exactly 0)
        de.close()


def test_parameters_outside_the_fp16_operand_range_take_the_cuda_core_contraction(oracle):
    """The UMMA operand elements are fp16: a catalogue whose log2e-scaled Gamma / mu leave their range is still sampled
    exactly in RG_MATH_TENSOR (the launch runs the CUDA-core contraction), and the tensor-core test hook refuses it."""
    import ctypes as C
    import torch
    from oracle import oracle as O
    from recogym_b200 import DeviceEnv, _lib as L
    cfg, sp = synthetic_env(300, 8, 21)
    sp["mu_organic"] = sp["mu_organic"].copy()
    sp["mu_organic"][5] = 3.0e4          # * log2(e) = 4.3e4 > 32 500
    sp["mu_organic"][9] = 3.0e4 - 1.0
    sp["mu_bandit"] = sp["mu_organic"] if cfg.number_of_flips == 0 else sp["mu_bandit"]
    oe = O.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                     cfg.sigma_omega_initial, cfg.sigma_omega, 21, cfg.change_omega_for_bandits)
    de = DeviceEnv(sp, cfg.sigma_omega_initial, cfg.sigma_omega, 21, cfg.change_omega_for_bandits)
    ro = oracle.rollout(oe, 0, 300)
    for mode in ("exact", "tensor"):
        r = de.rollout_device(300, "uniform", None, 0, 0, MODES[mode])
        assert np.array_equal(r["hash"].cpu().numpy().view(np.uint64), ro["hash"]), mode
        assert np.array_equal(r["last_view"].cpu().numpy(), ro["last_view"]) and int(r["totals"][9]) == 0
    views = set(np.unique(ro["last_view"]).tolist())
    assert views <= {5, 9} and len(views) == 2       # the two dominant products, one unit of logit apart
    om = torch.zeros((128, 8), dtype=torch.float64, device="cuda")
    sh = torch.zeros(128, dtype=torch.float32, device="cuda")
    out = torch.zeros((128, 384), dtype=torch.float32, device="cuda")
    rc = L.lib().rg_tc_logits_device(de._h, C.c_void_p(om.data_ptr()), C.c_void_p(sh.data_ptr()), C.c_void_p(out.data_ptr()), None)
    assert rc != 0 and b"operand" in L.lib().rg_last_error()
    de.close()


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_event_log_equals_oracle_and_reference_rows(oracle, mode):
    """The event-log output of the fused rollout == the rows of generate_logs: against the T2 reference logs
    (golden) and against the oracle's log at a larger size, every column, in the reference's row order."""
    from recogym_b200 import rollout_log
    for case in ("t2_seed7_P100_K5_u60_o5", "t2_seed5_P60_K8_u40_bandit_count_flips",
                 "t2_seed13_P40_K6_u60_o3_organic_user_count"):
        g = load_golden(case)
        de = device_env_from_golden(g)
        table = g["table"] if "table" in g.files else None
        r = rollout_log(de, int(g["n_users"]), str(g["policy"]), table, 0, int(g["n_organic_users"]), MODES[mode])
        for k in "utzvac":
            assert np.array_equal(r[k].astype(np.int64), g["log_" + k].astype(np.int64)), (case, k)
        de.close()
    oe, de = both_envs(300, 8, 5)
    table = np.random.RandomState(3).randint(0, 300, 300).astype(np.int32)
    ro = oracle.rollout(oe, 10, 400, policy="table", table=table, n_organic_users=7, log=True)
    r = rollout_log(de, 400, "table", table, 10, 7, MODES[mode], rows_per_user=20.0)   # forces the capacity retry
    for k in "utzvac":
        assert np.array_equal(r[k].astype(np.int64), ro["log"][k].astype(np.int64)), k
    de.close()


def test_fused_generate_logs_equals_stepwise_facade():
    """generate_logs(fused) and the per-step gym loop produce the same DataFrame (uniform logging policy and a
    BanditCount table agent, whose 'ps' column is its click-rate estimate)."""
    import recogym_b200 as rg
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 11, "num_products": 30, "K": 4})
    a = env.generate_logs(40, num_organic_offline_users=3)
    b = env.generate_logs(40, num_organic_offline_users=3, fused=False)
    agent = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": 30}))
    rs = np.random.RandomState(0)
    agent.pulls_a += rs.randint(0, 5, (30, 30)); agent.clicks_a += rs.randint(0, 2, (30, 30)) * (agent.pulls_a > 0)
    agent.ctr = (agent.clicks_a + 1) / (agent.pulls_a + 2)
    c = env.generate_logs(40, agent=agent)
    d = env.generate_logs(40, agent=agent, fused=False)
    # OrganicUserEventCounterAgent: device policy RG_POLICY_USER_VIEWS; 'ps' = views of a / all views so far
    uc = rg.agents.OrganicUserEventCounterAgent(rg.Configuration(
        {**rg.agents.organic_user_count_args, "num_products": 30, "random_seed": 11}))
    e = env.generate_logs(30, agent=uc)
    for u in (0, 7, 29):
        rows = e[e["u"] == u]
        seen = {}
        for _, r in rows.iterrows():
            if r["z"] == "organic":
                seen[int(r["v"])] = seen.get(int(r["v"]), 0) + 1
            else:
                assert int(r["a"]) in seen and r["ps"] == seen[int(r["a"])] / sum(seen.values())
    for x, y in ((a, b), (c, d)):
        assert list(x.columns) == list(y.columns) and len(x) == len(y)
        for col in ("t", "u", "v", "a", "c", "ps"):
            xv = x[col].to_numpy(dtype="float64", na_value=np.nan)
            yv = y[col].to_numpy(dtype="float64", na_value=np.nan)
            assert np.array_equal(np.isnan(xv), np.isnan(yv)), col
            assert np.array_equal(np.nan_to_num(xv, nan=-1), np.nan_to_num(yv, nan=-1)), col
        assert (x["z"] == y["z"]).all()

"""`reco-gym-v0`, the toy environment (reference: recogym/envs/reco_env_v0.py:17-69 on recogym/envs/abstract.py:90-327).

No latent state: the next organic view is uniform inside the current view's product cluster and a click is
Bernoulli(click_probs[action, view]); the clock never advances (every row of a log has t = 0).  The static parameters
(phi, click_probs, product_transition) are drawn and derived on the host with NumPy exactly as the reference does -- so they
equal it bit for bit for a given seed -- and handed to the CUDA library as search tables; every episode runs on the GPU
(csrc/rg_env0.cu, one thread per user).  The gym protocol (reset / step / step_offline) is served by REPLAYS: a user's
trajectory is a pure function of (seed, user, actions so far), so step(a) re-runs the current user with its action list
and returns the rows that are new.  There is no CPU path: without the library / a GPU the env raises.
"""
import ctypes as C

import numpy as np
from numpy.random.mtrand import RandomState
from scipy.special import expit as sigmoid

from .. import _lib as L
from .configuration import Configuration
from .context import DefaultContext
from .features.time import DefaultTimeGenerator
from .observation import Observation
from .reco_env_v1 import Discrete, env_args
from .session import OrganicSessions

organic, bandit, stop = 0, 1, 2

# reference: recogym/envs/reco_env_v0.py:5-13
env_0_args = {**env_args, "num_clusters": 2, "phi_var": 0.1}


def f(mat, offset=5):
    """Monotonic increasing squashing function (abstract.py:35-38)."""
    return sigmoid(mat - offset)


def choice_cdf(p):
    """the cumulative distribution RandomState.choice(n, p=p) searches (legacy mtrand: cumsum, divide by the last entry)"""
    cdf = np.asarray(p, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return cdf


def click_thresholds(click_probs):
    """rng.binomial(1, p) as a threshold on its single uniform U (numpy legacy random_binomial_inversion with n = 1;
    pinned against numpy in tests/test_env0_cpu.py): p <= 0.5: 1 iff U > exp(log(1 - p)); p > 0.5: the draw is made with
    1 - p and mirrored, 0 iff U > exp(log(1 - (1 - p))) -- stored NEGATED (always < 0 there: p > 0.5)."""
    p = np.asarray(click_probs, dtype=np.float64)
    low = p <= 0.5
    with np.errstate(divide="ignore"):
        thr = np.where(low, np.exp(np.log(1.0 - p)), -np.exp(np.log(1.0 - (1.0 - p))))
    return np.ascontiguousarray(thr)


def draw_static_params_v0(config, rng):
    """set_static_params (reco_env_v0.py:21-47) on the host, in the reference's operation order."""
    from numpy import array, eye, kron, ones, sqrt
    P, n_clusters = int(config.num_products), int(config.num_clusters)
    trans = array([
        [0, config.prob_organic_to_bandit, config.prob_leave_organic],
        [config.prob_bandit_to_organic, 0, config.prob_leave_organic],      # (sic: the reference uses prob_leave_organic here)
        [0.0, 0.0, 1.0],
    ])
    trans[0, 0] = 1 - sum(trans[0, :])
    trans[1, 1] = 1 - sum(trans[1, :])
    cluster_ratio = int(P / n_clusters)
    if cluster_ratio * n_clusters != P:
        raise ValueError("reco-gym-v0 needs num_products divisible by num_clusters (reco_env_v0.py:32-36 builds a "
                         "num_clusters x num_clusters block matrix)")
    ones_mat = ones((cluster_ratio, cluster_ratio))
    T = kron(eye(n_clusters), ones_mat)
    T = T / kron(T.sum(1), ones((P, 1))).T
    phi = rng.normal(scale=sqrt(config.phi_var), size=(P, P))
    click_probs = f(P / 5.0 * (T + T.T) + phi)
    return {"state_transition": trans, "product_transition": T, "phi": phi, "click_probs": click_probs,
            "initial_product_probs": ones((P)) / P}


def device_tables(sp):
    """search tables of the C ABI (include/recogym_b200.h rg_env0_desc) from the reference-shaped parameters"""
    T = sp["product_transition"]
    P = T.shape[0]
    rows, cluster_of = [], np.empty(P, np.int32)
    index = {}
    for v in range(P):
        key = T[v].tobytes()
        if key not in index:
            index[key] = len(rows)
            rows.append(choice_cdf(T[v]))
        cluster_of[v] = index[key]
    return {"init_cdf": np.ascontiguousarray(choice_cdf(sp["initial_product_probs"])),
            "view_cdf": np.ascontiguousarray(np.stack(rows)), "cluster_of": cluster_of,
            "click_thr": click_thresholds(sp["click_probs"]),
            "trans_cdf": np.ascontiguousarray(np.stack([choice_cdf(sp["state_transition"][s]) for s in range(3)]))}


class DeviceEnv0:
    """Handle of one rg_env0 (tables on the GPU)."""

    def __init__(self, tables, seed, device=0):
        self.P = int(tables["init_cdf"].shape[0])
        self._keep = {k: np.ascontiguousarray(v) for k, v in tables.items()}
        self.desc = self.make_desc(self._keep, seed, device)
        h = C.c_void_p()
        L.check(L.lib().rg_env0_create(C.byref(self.desc), C.byref(h)))
        self._h = h

    @staticmethod
    def make_desc(t, seed, device=0):
        d = L.Env0Desc()
        d.num_products, d.num_clusters = int(t["init_cdf"].shape[0]), int(t["view_cdf"].shape[0])
        d.init_cdf, d.view_cdf = t["init_cdf"].ctypes.data, t["view_cdf"].ctypes.data
        d.cluster_of, d.click_thr = t["cluster_of"].ctypes.data, t["click_thr"].ctypes.data
        d.trans_cdf = (C.c_double * 9)(*[float(x) for x in t["trans_cdf"].ravel()])
        d.seed, d.device = int(seed), int(device)
        return d

    def set_seed(self, seed):
        L.check(L.lib().rg_env0_set_seed(self._h, int(seed)))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().rg_env0_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rollout(self, n_users, policy="uniform", table=None, user_begin=0, n_organic_users=0, log=True, per_user=True,
                rows_per_user=110.0, _entry=None):
        """generate_logs(n_users, agent, n_organic_users) for users user_begin..: one launch.  -> dict of host arrays:
        totals[8], per-user counters, and the log rows in the reference's order (u, e, z, x, c, phantom)."""
        return run_rollout(lambda *a: (_entry or L.lib().rg_env0_rollout_host)(self._h, *a), self.P, n_users, policy, table,
                           user_begin, n_organic_users, log, per_user, rows_per_user)


def run_rollout(call, P, n_users, policy, table, user_begin, n_organic_users, log, per_user, rows_per_user):
    kind = {"uniform": L.POLICY0_UNIFORM, "table": L.POLICY0_TABLE, "replay": L.POLICY0_REPLAY}[policy]
    tab = np.ascontiguousarray(table, np.int32) if table is not None else np.zeros(0, np.int32)
    n = int(n_users) + int(n_organic_users)
    cap = int(n * rows_per_user) + 4096 if log else 0
    while True:
        out = L.Env0Out()
        res = {"totals": np.zeros(8, np.int64)}
        out.totals = res["totals"].ctypes.data
        if per_user:
            for k in ("n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "state"):
                res[k] = np.zeros(n, np.int32)
                setattr(out, k, res[k].ctypes.data)
        count = np.zeros(1, np.int64)
        rows = np.zeros((cap, 4), np.int32)
        if log:
            out.log, out.log_capacity, out.log_count = rows.ctypes.data, cap, count.ctypes.data
        L.check(call(kind, tab.ctypes.data if tab.size else None, int(tab.size), int(user_begin), int(n_organic_users),
                     int(n_users), C.byref(out)))
        if int(count[0]) <= cap:
            break
        cap = int(count[0]) + 1024      # the rows are a pure function of the inputs: run again with room
    if log:
        rows = rows[:int(count[0])]
        rows = rows[np.lexsort((rows[:, 1], rows[:, 0]))]       # (user, event index): the reference's row order
        flags = rows[:, 3]
        res.update({"u": rows[:, 0], "e": rows[:, 1], "z": (flags & 1).astype(np.int8), "x": rows[:, 2],
                    "c": ((flags >> 1) & 1).astype(np.int8), "phantom": ((flags >> 2) & 1).astype(np.int8)})
    return res


class RecoEnv0:
    """`gym.make('reco-gym-v0')` object of the reference (zero-arg constructor, then init_gym)."""

    metadata = {"render.modes": []}
    device_policies = ("uniform", "table")      # what rg_env0_rollout_* runs; other agents go through the gym protocol
    supports_device_training = False            # the count agents train through the gym protocol here (small catalogues)

    def __init__(self):
        self.first_step = True
        self.config = None
        self.state = None
        self.current_user_id = None
        self.current_time = None
        self.empty_sessions = OrganicSessions()
        self.agent = None
        self.device_env = None
        self._device = 0
        self._epoch_seed = None
        self._actions = []
        self._rows_seen = 0

    # ---------------------------------------------------------------------------------- set-up (abstract.py:59-88)
    def reset_random_seed(self, epoch=0):
        assert self.config.random_seed is not None
        self._epoch_seed = int(self.config.random_seed) + int(epoch)
        self.rng = RandomState(self._epoch_seed)
        if self.device_env is not None:
            self.device_env.set_seed(self._epoch_seed)

    def init_gym(self, args):
        self.config = Configuration(args)
        self.action_space = Discrete(self.config.num_products)
        self._device = int(args.get("device", 0)) if isinstance(args, dict) else 0
        self.time_generator = self.config.time_generator if "time_generator" in args else DefaultTimeGenerator(self.config)
        self.device_env = None
        self.reset_random_seed()
        self.agent = args["agent"] if "agent" in args else None
        self.set_static_params()
        self.reset_random_seed()

    def set_static_params(self):
        """reco_env_v0.py:21-47 (host NumPy, bit-identical to the reference) + upload of the search tables."""
        sp = draw_static_params_v0(self.config, self.rng)
        self.state_transition, self.product_transition = sp["state_transition"], sp["product_transition"]
        self.phi, self.click_probs = sp["phi"], sp["click_probs"]
        self.initial_product_probs = sp["initial_product_probs"]
        if self.device_env is not None:
            self.device_env.close()
        self.device_env = DeviceEnv0(device_tables(sp), self._epoch_seed, self._device)

    def __deepcopy__(self, memo):
        import copy
        new = RecoEnv0.__new__(RecoEnv0)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "device_env":
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new.device_env = None
        if self.device_env is not None:
            new.device_env = DeviceEnv0(self.device_env._keep, self._epoch_seed, self._device)
        return new

    def __getstate__(self):
        """picklable (the reference's drivers send envs to Pool workers: evaluate_agent.py:172-173): the device handle is
        dropped and re-created from the host tables on load"""
        st = dict(self.__dict__)
        de = st.pop("device_env", None)
        st["device_env"] = None
        st["_tables"] = de._keep if de is not None else None
        return st

    def __setstate__(self, st):
        tables = st.pop("_tables", None)
        self.__dict__.update(st)
        if tables is not None:
            self.device_env = DeviceEnv0(tables, self._epoch_seed, self._device)

    # ---------------------------------------------------------------------------------- gym protocol, by replay
    def reset(self, user_id=0):
        """abstract.py:90-103 + reco_env_v0.py:50-56"""
        self.first_step = True
        self.state = organic
        self.time_generator.reset()
        if self.agent:
            self.agent.reset()
        self.current_time = self.time_generator.new_time()
        self.current_user_id = user_id
        self.organic_views = np.zeros(self.config.num_products)
        self._actions = []
        self._rows_seen = 0
        self.product_view = None    # (the reference draws the initial view here; it is only read by update_product_view)

    def _replay(self):
        r = self.device_env.rollout(1, "replay", np.asarray(self._actions, np.int32), int(self.current_user_id), 0, log=True)
        new = slice(self._rows_seen, len(r["e"]))
        self._rows_seen = len(r["e"])
        self.state = stop if int(r["state"][0]) == 2 else bandit
        self.product_view = int(r["last_view"][0])
        return r, new

    def _sessions(self, r, rows):
        s = OrganicSessions()
        for z, x in zip(r["z"][rows], r["x"][rows]):
            if z == 0:
                s.next(DefaultContext(self.current_time, self.current_user_id), int(x))
        return s

    def step(self, action_id):
        """abstract.py:123-197"""
        if self.first_step:
            assert action_id is None
            self.first_step = False
            r, new = self._replay()
            return (Observation(DefaultContext(self.current_time, self.current_user_id), self._sessions(r, new)), None,
                    self.state == stop, {})
        assert action_id is not None
        self._actions.append(int(action_id))
        r, new = self._replay()
        first = new.start
        reward = int(r["c"][first])         # the first new row is the bandit event of this action
        return (Observation(DefaultContext(self.current_time, self.current_user_id),
                            self._sessions(r, slice(first + 1, new.stop))), reward, self.state == stop, {})

    def step_offline(self, observation, reward, done):
        """abstract.py:199-239"""
        if self.first_step:
            action = None
        else:
            assert observation is not None
            if self.agent:
                action = self.agent.act(observation, reward, done)
            else:
                # the default uniform logging policy (abstract.py:214), drawn from the policy stream of the counter RNG
                # at the user's current event index (= the rows of its log so far)
                P = self.config.num_products
                a = L.lib().rg_spec_policy_uniform_action(self._epoch_seed, int(self.current_user_id), int(self._rows_seen), P)
                action = {"t": observation.context().time(), "u": observation.context().user(), "a": int(a),
                          "ps": 1.0 / P, "ps-a": (np.ones(P) / P if self.config.with_ps_all else ())}
        if done:
            return (action, Observation(DefaultContext(self.current_time, self.current_user_id), self.empty_sessions),
                    0, done, None)
        observation, reward, done, info = self.step(action["a"] if action is not None else None)
        return action, observation, reward, done, info

    # ---------------------------------------------------------------------------------- logs
    def generate_logs(self, num_offline_users, agent=None, num_organic_offline_users=0, user_begin=0):
        """abstract.py:241-327 in ONE launch: the uniform logging policy, or an agent that compiles to a last-view table
        (`device_policy()` -> {'policy': 'table', 'table': ...}); any other agent goes through the gym protocol above."""
        import pandas as pd
        pol = {"policy": "uniform"}
        agent = agent or self.agent
        if agent is not None:
            pol = agent.device_policy() if hasattr(agent, "device_policy") else None
            if pol is None or pol.get("policy") != "table":
                return self._generate_logs_stepwise(num_offline_users, agent, num_organic_offline_users, user_begin)
        table = pol.get("table")
        if table is not None and hasattr(table, "is_cuda"):
            table = table.detach().cpu().numpy()
        r = self.device_env.rollout(num_offline_users, pol["policy"], table, int(user_begin), num_organic_offline_users)
        org = r["z"] == 0
        P = self.config.num_products
        ps = np.where(org, np.nan, 1.0 / P if pol["policy"] == "uniform" else 1.0)
        ps_table = pol.get("ps_table")
        if ps_table is not None:        # e.g. BanditCount: 'ps' = its click-rate estimate at the acting view
            lv = pd.Series(np.where(org, r["x"], np.nan)).groupby(r["u"]).ffill().to_numpy()
            b = np.flatnonzero(~org)
            ps[b] = np.asarray(ps_table.detach().cpu().numpy() if hasattr(ps_table, "is_cuda") else ps_table)[lv[b].astype(np.int64)]
        n = len(org)
        psa = np.empty(n, dtype=object)
        psa[:] = None
        empty = np.empty(1, dtype=object)
        empty[0] = ()
        psa[~org] = empty
        return pd.DataFrame({
            "t": np.zeros(n, np.float32),                               # the v0 clock never advances
            "u": pd.arrays.IntegerArray(r["u"].astype(np.int32), np.zeros(n, bool)),
            "z": np.where(org, "organic", "bandit"),
            "v": pd.arrays.IntegerArray(r["x"].astype(np.int32), ~org),
            "a": pd.arrays.IntegerArray(r["x"].astype(np.int32), org),
            "c": np.where(org, np.float32("nan"), r["c"].astype(np.float32)),
            "ps": ps, "ps-a": psa})

    def generate_logs_counters(self, num_offline_users, policy="uniform", table=None, num_organic_offline_users=0,
                               user_begin=0, epsilon=0.0, **_):
        """generate_logs() reduced to its counters, all users in one launch: successes / failures as test_agent and
        verify_agents count them (bench_agents.py:203-206, evaluate_agent.py:742-745: phantom rows are failures)."""
        if policy not in self.device_policies or epsilon:
            raise NotImplementedError("reco-gym-v0 runs the uniform and table policies on the device; got %r" % (policy,))
        if table is not None and hasattr(table, "is_cuda"):
            table = table.detach().cpu().numpy()
        tot = self.device_env.rollout(num_offline_users, policy, table, int(user_begin), num_organic_offline_users,
                                      log=False, per_user=False)["totals"]
        return {"successes": int(tot[4]), "failures": int(tot[3] + tot[5] - tot[4]), "totals": tot}

    def _generate_logs_stepwise(self, num_offline_users, agent, num_organic_offline_users, user_begin=0):
        import pandas as pd
        old_agent, self.agent = self.agent, agent
        data = {k: [] for k in ("t", "u", "z", "v", "a", "c", "ps", "ps-a")}

        def store_organic(observation):
            for s in observation.sessions():
                data["t"].append(s["t"]); data["u"].append(s["u"]); data["z"].append("organic")
                data["v"].append(s["v"]); data["a"].append(None); data["c"].append(None)
                data["ps"].append(None); data["ps-a"].append(None)

        def store_bandit(action, reward):
            if action:
                data["t"].append(action["t"]); data["u"].append(action["u"]); data["z"].append("bandit")
                data["v"].append(None); data["a"].append(action["a"]); data["c"].append(reward)
                data["ps"].append(action["ps"]); data["ps-a"].append(action["ps-a"] if "ps-a" in action else ())

        uid = int(user_begin)
        for _ in range(num_organic_offline_users):
            self.reset(uid); uid += 1
            observation, _, _, _ = self.step(None)
            store_organic(observation)
        for _ in range(num_offline_users):
            self.reset(uid); uid += 1
            observation, reward, done, _ = self.step(None)
            while not done:
                store_organic(observation)
                action, observation, reward, done, _ = self.step_offline(observation, reward, done)
                store_bandit(action, reward)
            store_organic(observation)
            action, _, reward, done, _ = self.step_offline(observation, reward, done)
            assert done, "Done must not be changed!"
            store_bandit(action, reward)
        self.agent = old_agent
        data["t"] = np.array(data["t"], dtype=np.float32)
        for k in ("u", "v", "a"):
            data[k] = pd.array(data[k], dtype=pd.Int32Dtype())
        data["c"] = np.array(data["c"], dtype=np.float32)
        return pd.DataFrame().from_dict(data)

"""RecoEnv1: the reference's gym-style environment object, backed by the CUDA kernels.

Mirrors AbstractEnv + RecoEnv1 (reference: recogym/envs/abstract.py:46-327,
recogym/envs/reco_env_v1.py:18-174): same constructor, `init_gym(args)`, `reset(user_id)`,
`step(action_id)`, `step_offline(obs, reward, done)`, `generate_logs(...)`, `reset_random_seed(epoch)`,
same readable attributes and the same Observation / action-dict shapes, so the reference's Python
agents run unchanged.  One user at a time through this object is the compatibility path; the
throughput path is `DeviceEnv.rollout_device` / `BatchState` (millions of users per launch), which
`generate_logs_counters`, `recogym_b200.test_agent` and `recogym_b200.verify_agents` use.

Dynamics are driven by the counter RNG of DESIGN.md (Philox keyed by user id and event time), not
by the reference's sequential MT19937 stream; static parameters are bit-identical to the
reference's for the same `random_seed`.
"""
from copy import deepcopy

import numpy as np
from numpy.random.mtrand import RandomState

from .. import _lib as L
from .configuration import Configuration
from .context import DefaultContext
from .device_env import BatchState, DeviceEnv
from .features.time import DefaultTimeGenerator, NormalTimeGenerator
from .observation import Observation
from .session import OrganicSessions
from .static_params import draw_static_params, make_beta

organic, bandit, stop = 0, 1, 2

# Arguments shared between all environments (reference: recogym/envs/abstract.py:20-31).
env_args = {
    "num_products": 10,
    "num_users": 100,
    "random_seed": np.random.randint(2 ** 31 - 1),
    "prob_leave_bandit": 0.01,
    "prob_leave_organic": 0.01,
    "prob_bandit_to_organic": 0.05,
    "prob_organic_to_bandit": 0.25,
    "normalize_beta": False,
    "with_ps_all": False,
}

# reference: recogym/envs/reco_env_v1.py:18-29.  `num_clusters`/`phi_var` are inherited in the
# reference only because importing reco_env_v0 mutates the shared dict (reco_env_v0.py:7-13);
# bench_agents' cache key reads them (bench_agents.py:29,33), so they are kept.
env_1_args = {
    **env_args,
    "num_clusters": 2,
    "phi_var": 0.1,
    "K": 5,
    "sigma_omega_initial": 1,
    "sigma_omega": 0.1,
    "number_of_flips": 0,
    "sigma_mu_organic": 3,
    "change_omega_for_bandits": False,
    "normalize_beta": False,
}


class Discrete:
    """Stand-in for gym.spaces.Discrete (abstract.py:69): `n` and `sample()`."""

    def __init__(self, n):
        self.n = int(n)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class RecoEnv1:
    """`gym.make('reco-gym-v1')` object of the reference (zero-arg constructor, then init_gym)."""

    metadata = {"render.modes": []}

    def __init__(self):
        self.first_step = True
        self.config = None
        self.state = None
        self.current_user_id = None
        self.current_time = None
        self.empty_sessions = OrganicSessions()
        self.agent = None
        self.device_env = None
        self._batch = None
        self._device = 0
        self._epoch_seed = None
        self._user_key_offset = 0   # RNG key of the facade's current user = its label + this (generate_logs(user_begin=))

    # ---------------------------------------------------------------------------------- set-up
    def reset_random_seed(self, epoch=0):
        """abstract.py:59-62.  `rng` stays a RandomState for API compatibility (static parameters
        are drawn from it); the dynamics use the device counter RNG keyed by the same seed."""
        assert self.config.random_seed is not None
        self._epoch_seed = int(self.config.random_seed) + int(epoch)
        self.rng = RandomState(self._epoch_seed)
        if self.device_env is not None:
            self.device_env.set_seed(self._epoch_seed)

    def init_gym(self, args):
        """abstract.py:64-88.  Accepts every key of env_1_args plus optional `agent`,
        `time_generator`, and the new optional `device` (CUDA ordinal, default 0)."""
        self.config = Configuration(args)
        self.action_space = Discrete(self.config.num_products)
        self._device = int(args.get("device", 0)) if isinstance(args, dict) else 0
        if "time_generator" not in args:
            self.time_generator = DefaultTimeGenerator(self.config)
        else:
            self.time_generator = self.config.time_generator
            if not isinstance(self.time_generator, (DefaultTimeGenerator, NormalTimeGenerator)):
                raise NotImplementedError(
                    "recogym_b200 implements DefaultTimeGenerator and NormalTimeGenerator (reference: "
                    "recogym/envs/features/time/); got %r" % type(self.time_generator).__name__)
        self.device_env = None
        self.reset_random_seed()
        self.agent = args["agent"] if "agent" in args else None
        self.set_static_params()
        self.reset_random_seed()

    def set_static_params(self):
        """reco_env_v1.py:51-75 (host NumPy, bit-identical to the reference) + upload to the GPU."""
        sp = draw_static_params(self.config, self.rng)
        self.state_transition = sp["state_transition"]
        self.Gamma, self.mu_organic = sp["Gamma"], sp["mu_organic"]
        self.beta, self.mu_bandit = sp["beta"], sp["mu_bandit"]
        self._upload()

    def generate_beta(self, number_of_flips):
        """reco_env_v1.py:133-174; callable from outside (notebooks do), re-uploads parameters."""
        self.beta, self.mu_bandit = make_beta(self.Gamma, self.mu_organic, number_of_flips,
                                              self.config.normalize_beta)
        self._upload()

    def normalize_beta(self):
        self.beta = self.beta / np.sqrt((self.beta ** 2).sum(1)[:, np.newaxis])
        self._upload()

    def _upload(self):
        if self.device_env is not None:
            self.device_env.close()
        c = self.config
        self.device_env = DeviceEnv(
            {"Gamma": self.Gamma, "mu_organic": self.mu_organic, "beta": self.beta, "mu_bandit": self.mu_bandit,
             "state_transition": self.state_transition},
            c.sigma_omega_initial, c.sigma_omega, self._epoch_seed, c.change_omega_for_bandits, self._device,
            self._normal_time())
        self._batch = None

    def _normal_time(self):
        """(mu, sigma) when the env's clock is a NormalTimeGenerator, else None (event index == clock)."""
        tg = getattr(self, "time_generator", None)
        if isinstance(tg, NormalTimeGenerator):
            return float(tg.normal_time_mu), float(tg.normal_time_sigma)
        return None

    def _clock(self, event_index):
        """current_time at the given event index of the facade's current user.  Default clock: the index itself.
        NormalTimeGenerator: the sequential binary64 sum of the steps of the calls 0 .. index-1 (the kernels' Philox
        slot 2 at (user key, call)), as normal_time_generator.py:25-28 accumulates them."""
        nt = self._normal_time()
        if nt is None:
            return int(event_index)
        cum = self._clock_cum
        key = int(self.current_user_id) + int(getattr(self, "_user_key_offset", 0))
        while len(cum) <= event_index:
            cum.append(cum[-1] + L.lib().rg_spec_time_step(self._epoch_seed, key, len(cum) - 1, nt[0], nt[1]))
        return cum[event_index]

    # ---------------------------------------------------------------------------- copy / pickle
    def __deepcopy__(self, memo):
        new = RecoEnv1.__new__(RecoEnv1)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("device_env", "_batch"):
                continue
            new.__dict__[k] = deepcopy(v, memo)
        new.device_env = self.device_env.clone() if self.device_env is not None else None
        new._batch = None
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["device_env"] = None
        st["_batch"] = None
        st["_needs_upload"] = self.device_env is not None
        return st

    def __setstate__(self, st):
        needs = st.pop("_needs_upload", False)
        self.__dict__.update(st)
        if needs:
            self._upload()

    # -------------------------------------------------------------------------------- episode
    @property
    def omega(self):
        """(K,1) latent state of the current user (host copy)."""
        if self._batch is None:
            return None
        return self._batch.omega.cpu().numpy().reshape(-1, 1)

    def reset(self, user_id=0):
        """abstract.py:90-103 + reco_env_v1.py:78-82.  The first organic session is produced by the
        same launch and handed out by step(None)."""
        self.first_step = True
        self.state = organic
        self.time_generator.reset()
        if self.agent:
            self.agent.reset()
        self.time_generator.new_time()
        self.current_user_id = user_id
        self._event_index = 0            # Markov events of this user so far == the RNG's event time
        self._clock_cum = [0.0]
        self.current_time = self._clock(0)
        self.organic_views = np.zeros(self.config.num_products)
        key = int(user_id) + int(getattr(self, "_user_key_offset", 0))
        if self._batch is None:
            self._batch = BatchState(self.device_env, [key], sess_cap=64)
        else:
            self._batch.user_id.fill_(key)
        self._pending_reset = True

    def _collect_sessions(self, h, t_before):
        """Turn the device session buffer (plus continuation launches) into OrganicSessions."""
        sessions = OrganicSessions()
        t = t_before
        while True:
            n = int(h["sess_len"][0])
            for v in h["sess_v"][0, :n]:
                sessions.next(DefaultContext(self._clock(t), self.current_user_id), int(v))
                t += 1
            if int(h["state"][0]) != organic:
                break
            self._batch.continue_sessions(L.MATH_EXACT)
            h = self._batch.read()
        self.state = int(h["state"][0])
        self._set_event_index(int(h["t"][0]))
        return sessions

    def _set_event_index(self, idx):
        self._event_index = idx
        self.current_time = self._clock(idx)
        self.time_generator.current_time = self._clock(idx + 1)   # the generator's own state after new_time()

    def step(self, action_id):
        """abstract.py:123-197: returns (Observation, reward, done, info)."""
        info = {}
        if self.first_step:
            assert action_id is None
            self.first_step = False
            self._batch.reset(L.MATH_EXACT)
            self._pending_reset = False
            sessions = self._collect_sessions(self._batch.read(), 0)
            return (Observation(DefaultContext(self.current_time, self.current_user_id), sessions), None,
                    self.state == stop, info)
        assert action_id is not None
        a = int(action_id)
        assert 0 <= a < self.config.num_products
        t_before = self._event_index
        self._batch.step(np.array([a], np.int32), L.MATH_EXACT)
        h = self._batch.read()
        reward = int(h["reward"][0])
        if int(h["sess_len"][0]) > 0 or int(h["state"][0]) == organic:
            sessions = self._collect_sessions(h, t_before + 1)
        else:
            sessions = self.empty_sessions
            self.state = int(h["state"][0])
            self._set_event_index(int(h["t"][0]))
        return (Observation(DefaultContext(self.current_time, self.current_user_id), sessions), reward,
                self.state == stop, info)

    def step_offline(self, observation, reward, done):
        """abstract.py:199-239: act with the env's agent, or with the default uniform logging
        policy (drawn from the policy stream of the counter RNG at the current event time)."""
        if self.first_step:
            action = None
        else:
            assert observation is not None
            if self.agent:
                action = self.agent.act(observation, reward, done)
            else:
                P = self.config.num_products
                a = L.lib().rg_spec_policy_uniform_action(
                    self._epoch_seed, int(self.current_user_id) + int(getattr(self, "_user_key_offset", 0)),
                    int(self._event_index), P)
                action = {
                    "t": observation.context().time(),
                    "u": observation.context().user(),
                    "a": int(a),
                    "ps": 1.0 / P,
                    "ps-a": (np.ones(P) / P if self.config.with_ps_all else ()),
                }
        if done:
            return (action, Observation(DefaultContext(self.current_time, self.current_user_id), self.empty_sessions),
                    0, done, None)
        observation, reward, done, info = self.step(action["a"] if action is not None else None)
        return action, observation, reward, done, info

    def generate_logs(self, num_offline_users, agent=None, num_organic_offline_users=0, fused=None, user_begin=0):
        """abstract.py:241-327: DataFrame[t,u,z,v,a,c,ps,ps-a] of a rollout.  u/v/a are Int32 (the reference's
        UInt16 overflows at 65 535 users / int16 at 32 767 products).

        user_begin: the reference env draws from ONE sequential stream, so every call sees fresh users; here a user is
        the counter-RNG key `user_begin + u` (u = the 0-based label in the log), and a call is reproducible: the default
        replays users 0..N-1.  Drivers that train and then evaluate pass the first unused key (test_agent, verify_agents).

        fused=None (default): when the acting policy runs on the device (no agent = the default uniform logging
        policy, or an agent with `device_policy()`), all users are simulated by ONE fused CUDA rollout whose
        event log is sorted into the reference's row order -- same rows as the per-step loop, ~10^4x faster.
        Any other Python agent (or fused=False, or with_ps_all) goes one user after the other through
        step()/step_offline()."""
        import pandas as pd
        from ..policy import device_policy
        pol = device_policy(agent if agent else self.agent, self.config.num_products)
        if fused is None:
            fused = pol is not None and not getattr(self.config, "with_ps_all", False)
        if fused:
            if pol is None:
                raise ValueError("fused generate_logs needs a device policy (no agent, or agent.device_policy())")
            return self._generate_logs_fused(num_offline_users, pol, num_organic_offline_users, user_begin)
        self._user_key_offset = int(user_begin)
        if agent:
            old_agent = self.agent
            self.agent = agent
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

        uid = 0
        for _ in range(num_organic_offline_users):
            self.reset(uid)
            uid += 1
            observation, _, _, _ = self.step(None)
            store_organic(observation)
        for _ in range(num_offline_users):
            self.reset(uid)
            uid += 1
            observation, reward, done, _ = self.step(None)
            while not done:
                store_organic(observation)
                action, observation, reward, done, _ = self.step_offline(observation, reward, done)
                store_bandit(action, reward)
            store_organic(observation)
            action, _, reward, done, _ = self.step_offline(observation, reward, done)
            assert done, "Done must not be changed!"
            store_bandit(action, reward)

        data["t"] = np.array(data["t"], dtype=np.float32)
        data["u"] = pd.array(data["u"], dtype=pd.Int32Dtype())
        data["v"] = pd.array(data["v"], dtype=pd.Int32Dtype())
        data["a"] = pd.array(data["a"], dtype=pd.Int32Dtype())
        data["c"] = np.array(data["c"], dtype=np.float32)
        if agent:
            self.agent = old_agent
        self._user_key_offset = 0
        return pd.DataFrame().from_dict(data)

    def _generate_logs_fused(self, num_offline_users, pol, num_organic_offline_users, user_begin=0):
        """ONE fused rollout with the event-log output, ordered and turned into the DataFrame's columns on the device
        (rg_generate_logs_device: scan + scatter, t / u / z / v / a / c / ps incl. the NormalTimeGenerator clock; the 'ps' of
        OrganicUserEventCounterAgent's actions by logs.user_views_ps), one
        device -> host copy per column, DataFrame built without per-row Python objects (logs.DeviceLog.to_dataframe)."""
        from ..logs import generate_logs_device
        log = generate_logs_device(self.device_env, num_offline_users, pol["policy"], pol.get("table"), pol.get("ps_table"),
                                   int(user_begin), num_organic_offline_users, epsilon=pol.get("epsilon", 0.0))
        return log.to_dataframe(zero_based_users=True)      # labels are 0-based per call, as in the reference

    # --------------------------------------------------------------------- throughput entry point
    def generate_logs_counters(self, num_offline_users, policy="uniform", table=None, num_organic_offline_users=0,
                               user_begin=0, math_mode=L.MATH_FAST, hist=False, per_user=False, epsilon=0.0):
        """generate_logs() reduced to its counters, all users in ONE fused launch.  Returns a dict of
        host numpy arrays; successes/failures as test_agent counts them (phantom rows are failures,
        bench_agents.py:203-206)."""
        r = self.device_env.rollout_device(num_offline_users, policy, table, user_begin, num_organic_offline_users,
                                           math_mode, per_user=per_user, hist=hist, epsilon=epsilon)
        out = {k: v.cpu().numpy() for k, v in r.items() if not k.startswith("_")}
        tot = out["totals"]
        out["successes"] = int(tot[4])
        out["failures"] = int(tot[3] + tot[5] - tot[4])
        return out

"""Runs the UNMODIFIED reference (criteo-research/reco-gym at /root/reference) -- TEST INFRASTRUCTURE ONLY.

Two oracle tiers are produced from the reference's own code:

  T1  literal reference: MT19937 RandomState, exactly what a user of reco-gym gets.
      Pins the static parameters (Gamma, mu_organic, beta, state_transition) bit-for-bit and
      pins distributions / the 'Getting Started.ipynb' seed-42 trace.
  T2  reference logic + injected counter RNG: the reference's step()/update_state()/
      draw_click()/update_product_view()/generate_logs() run UNMODIFIED, but `env.rng` is a
      PhiloxUserRNG whose draws are keyed by (seed, user, event time, slot) -- the RNG spec of
      DESIGN.md.  Its logs are what oracle/rg_oracle.c and the CUDA kernels must reproduce
      bit-for-bit in every discrete column.

/root/reference only exists in the build container; nothing that runs on the GPU box imports this
module (the goldens it produces are committed under tests/golden/).  Three import shims are needed
and no reference source is edited: a fake `gym`, an empty `matplotlib(.pyplot)`, `scipy.rand`.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RECOGYM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "recogym"))


def _install_shims():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # gym.Env stand-in (abstract.py:46-49 only calls __init__)
            def __init__(self):
                pass

        class Discrete:
            def __init__(self, n):
                self.n = n

        _registry = {}

        def register(id, entry_point):
            _registry[id] = entry_point

        def make(id):
            mod, cls = _registry[id].split(":")
            import importlib
            return getattr(importlib.import_module(mod), cls)()

        gym.Env = Env
        gym.make = make
        spaces = types.ModuleType("gym.spaces")
        spaces.Discrete = Discrete
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.register = register
        envs.registration = registration
        gym.spaces, gym.envs = spaces, envs
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                            "gym.envs.registration": registration})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    import scipy
    if not hasattr(scipy, "rand"):
        scipy.rand = np.random.rand


_REF = None


def load_reference():
    """import recogym from /root/reference under the shims; returns the module."""
    global _REF
    if _REF is None:
        if not reference_available():
            raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
        _install_shims()
        import tqdm
        import tqdm.std

        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import recogym  # noqa: the reference package
        import recogym.envs.abstract as ab

        # silence progress bars (display only; abstract.py:293,299)
        ab.trange = lambda n, **kw: range(n)
        _REF = recogym
    return _REF


# ------------------------------------------------------------------------------------------ T2 RNG
class PhiloxUserRNG:
    """Drop-in for `env.rng` (numpy RandomState) whose draws follow the build's RNG spec.

    Call sites in the reference and the stream each one maps to:
      reset                reco_env_v1.py:80   normal(0, s0, (K,1))      -> normals(user, t=0)
      update_product_view  reco_env_v1.py:124  choice(P, p)              -> event uniform 0 at t
      update_state         reco_env_v1.py:87   choice(3, p)              -> event uniform 1 at t
      update_state         reco_env_v1.py:96   normal(omega, s, (K,1))   -> normals(user, t) (t already advanced)
      draw_click           reco_env_v1.py:112  choice([0,1], p)          -> event uniform 0 at t
      step_offline         abstract.py:214     choice(P)                 -> policy stream at t
    The caller is identified by its function name (frame inspection): the transforms themselves
    (inverse CDF with searchsorted-right, loc+scale*z) are numpy's legacy ones.
    """

    def __init__(self, env, seed):
        self.env = env
        self.seed = int(seed)

    @property
    def O(self):
        from oracle import oracle as O  # not stored: the env must stay deep-copyable
        return O

    def _ut(self):
        # streams are keyed by the EVENT INDEX: the clock value under the default clock, the number of new_time() calls
        # since reset() minus one under the (float) NormalTimeGenerator clock
        trng = getattr(self.env.time_generator, "rng", None)
        if isinstance(trng, PhiloxTimeRNG):
            return int(self.env.current_user_id), trng.calls - 1
        return int(self.env.current_user_id), int(self.env.current_time)

    def normal(self, loc=0.0, scale=1.0, size=None):
        caller = sys._getframe(1).f_code.co_name
        assert caller in ("reset", "update_state"), caller
        u, t = self._ut()
        n = int(np.prod(size))
        z = self.O.normals(self.seed, u, t, n).reshape(size)
        return loc + scale * z

    def choice(self, a, p=None):
        caller = sys._getframe(1).f_code.co_name
        u, t = self._ut()
        if caller == "step_offline":
            assert p is None
            return self.O.policy_uniform_action(self.seed, u, t, int(a))
        which = {"update_product_view": 0, "draw_click": 0, "update_state": 1}[caller]
        x = self.O.event_uniform(self.seed, u, t, which)
        p = np.asarray(p, dtype=np.float64)
        cdf = p.cumsum()
        cdf /= cdf[-1]
        idx = int(cdf.searchsorted(x, side="right"))
        return idx if np.isscalar(a) or isinstance(a, (int, np.integer)) else a[idx]


class PhiloxTimeRNG:
    """Drop-in for NormalTimeGenerator.rng (normal_time_generator.py:23,27): the i-th new_time() call since reset() of
    user u draws mu + sigma * z(seed, u, i, time slot)."""

    def __init__(self, seed):
        self.seed = int(seed)
        self.user = 0
        self.calls = 0

    def normal(self, loc=0.0, scale=1.0):
        from oracle import oracle as O
        z = O.time_normal(self.seed, self.user, self.calls)
        self.calls += 1
        return loc + scale * z


class PhiloxAgentRNG:
    """Drop-in for the `rng` of an agent's model (e.g. OrganicUserEventCounterModel.rng,
    agents/organic_user_count.py:43): draws come from the POLICY stream of the RNG spec at the env's
    current (user, time) -- first uniform for the explore coin `choice([True, False], p)`
    (organic_user_count.py:49-52), second uniform for the action `choice(P, p)` (:65)."""

    def __init__(self, env, seed):
        self.env = env
        self.seed = int(seed)

    def __deepcopy__(self, memo):
        # evaluate_agent deep-copies the agent at every step boundary (evaluate_agent.py:59,84,139-142): the copy must keep
        # reading the (user, time) of the ONE env that is being stepped, not of a copy of it
        return PhiloxAgentRNG(self.env, self.seed)

    def choice(self, a, p=None):
        from oracle import oracle as O
        u, t = self.env.rng._ut() if isinstance(self.env.rng, PhiloxUserRNG) else (int(self.env.current_user_id),
                                                                                     int(self.env.current_time))
        is_coin = not (np.isscalar(a) or isinstance(a, (int, np.integer)))
        x = O.policy_uniform53(self.seed, u, t, 0 if is_coin else 1)
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        idx = int(cdf.searchsorted(x, side="right"))
        return a[idx] if is_coin else idx


def make_env(args, tier, normal_time=None):
    """Build the reference RecoEnv1 with `args` merged over env_1_args. tier: 'T1' | 'T2'.
    normal_time=(mu, sigma): the reference's NormalTimeGenerator as the env's clock (T2: its rng re-keyed)."""
    ref = load_reference()
    from copy import deepcopy
    env = ref.envs.reco_env_v1.RecoEnv1()
    full = {**deepcopy(ref.env_1_args), **args}
    if normal_time is not None:
        full["normal_time_mu"], full["normal_time_sigma"] = normal_time
        full["time_generator"] = ref.NormalTimeGenerator(ref.Configuration(full))
    env.init_gym(full)
    if normal_time is not None and tier == "T2":
        tg = env.time_generator
        tg.rng = PhiloxTimeRNG(full["random_seed"])
        tg_reset, env_reset = tg.reset, env.reset

        def reset_clock(_tg=tg, _orig=tg_reset):      # time_generator.reset() also restarts the call counter
            _orig()
            _tg.rng.calls = 0

        def reset_env(user_id=0, _tg=tg, _orig=env_reset):
            _tg.rng.user = int(user_id)               # AbstractEnv.reset draws the first increment BEFORE it stores user_id
            return _orig(user_id)

        tg.reset = reset_clock
        env.reset = reset_env
    if tier == "T2":
        # static parameters stay the literal reference's (drawn from RandomState(seed) in
        # init_gym); only the dynamics stream is replaced.  reset_random_seed(epoch) re-keys.
        def reset_random_seed(epoch=0, _env=env):
            _env.rng = PhiloxUserRNG(_env, _env.config.random_seed + epoch)
        env.reset_random_seed = reset_random_seed
        env.reset_random_seed()
    return env


def static_params(env):
    return {
        "Gamma": np.array(env.Gamma, np.float64),
        "mu_organic": np.array(env.mu_organic, np.float64).ravel(),
        "beta": np.array(env.beta, np.float64),
        "mu_bandit": np.array(env.mu_bandit, np.float64).ravel(),
        "state_transition": np.array(env.state_transition, np.float64),
    }


def logs_to_arrays(df):
    """generate_logs DataFrame -> integer columns (v/a/c = -1 where the reference has NaN)."""
    z = (df["z"].to_numpy() == "bandit").astype(np.int8)
    def col(name):
        x = df[name].to_numpy(dtype="float64", na_value=np.nan)
        return np.where(np.isnan(x), -1, x).astype(np.int64)
    return {"u": col("u").astype(np.int32), "t": df["t"].to_numpy().astype(np.int32), "z": z,
            "v": col("v").astype(np.int32), "a": col("a").astype(np.int32), "c": col("c").astype(np.int8)}


# ------------------------------------------------------------------------------------------ reco-gym-v0
class PhiloxUserRNGv0:
    """Drop-in for `env.rng` of the reference's RecoEnv0 (reco_env_v0.py) after init_gym: the dynamics draw from the
    build's RNG spec for v0 (include/recogym_b200.h "reco-gym-v0"), keyed by the user and its EVENT INDEX e = the number
    of update_state() calls so far (the v0 clock never advances, so the time cannot serve as the key):
      reset                reco_env_v0.py:54   choice(P, p=initial_product_probs)   -> slot 3, counter 0, uniform 0
      update_product_view  reco_env_v0.py:68   choice(P, p=product_transition[v])   -> event uniform 0 at e
      draw_click           reco_env_v0.py:65   binomial(1, p)                       -> event uniform 0 at e (numpy's legacy
                                                                                       inversion for n = 1 on that uniform)
      update_state         reco_env_v0.py:60   choice(3, p)                         -> event uniform 1 at e; then e += 1
      step_offline         abstract.py:214     choice(P)                            -> policy stream at e
    """

    def __init__(self, env, seed):
        self.env = env
        self.seed = int(seed)
        self.e = 0

    @staticmethod
    def _search(p, x):
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        return int(cdf.searchsorted(x, side="right"))

    def choice(self, a, p=None):
        from oracle import oracle as O
        caller = sys._getframe(1).f_code.co_name
        u = int(self.env.current_user_id)
        if caller == "reset":
            self.e = 0
            w = O.philox(self.seed, u, 0, 3)
            x = ((int(w[0]) >> 5) * 67108864.0 + (int(w[1]) >> 6)) / 9007199254740992.0
            return self._search(p, x)
        if caller == "step_offline":
            assert p is None
            return O.policy_uniform_action(self.seed, u, self.e, int(a))
        if caller == "update_product_view":
            return self._search(p, O.event_uniform(self.seed, u, self.e, 0))
        assert caller == "update_state", caller
        s = self._search(p, O.event_uniform(self.seed, u, self.e, 1))
        self.e += 1
        return s

    def binomial(self, n, p):
        from oracle import oracle as O
        from oracle.v0_oracle import binomial1
        assert n == 1 and sys._getframe(1).f_code.co_name == "draw_click"
        return binomial1(float(p), O.event_uniform(self.seed, int(self.env.current_user_id), self.e, 0))


def make_env0(args, tier):
    """The reference's RecoEnv0 with `args` over env_0_args.  T1: literal (MT19937); T2: static parameters from the literal
    init_gym, dynamics from PhiloxUserRNGv0."""
    ref = load_reference()
    from copy import deepcopy
    import recogym.envs.reco_env_v0 as r0
    env = r0.RecoEnv0()
    full = {**deepcopy(r0.env_0_args), "num_clusters": 2, "phi_var": 0.1, **args}
    env.init_gym(full)
    if tier == "T2":
        def reset_random_seed(epoch=0, _env=env):
            _env.rng = PhiloxUserRNGv0(_env, _env.config.random_seed + epoch)
        env.reset_random_seed = reset_random_seed
        env.reset_random_seed()
    return env


def static_params_v0(env):
    return {"state_transition": np.array(env.state_transition, np.float64),
            "product_transition": np.array(env.product_transition, np.float64), "phi": np.array(env.phi, np.float64),
            "click_probs": np.array(env.click_probs, np.float64),
            "initial_product_probs": np.array(env.initial_product_probs, np.float64)}

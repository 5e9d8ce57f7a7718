"""ctypes wrapper of oracle/librg_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (recogym_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile the C restatement (gcc, seconds). Building the checker is not using it."""
    so = os.path.join(_HERE, "librg_oracle.so")
    src = os.path.join(_HERE, "rg_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return so


class _Env(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("K", C.c_int32),
        ("Gamma", C.c_void_p), ("mu_organic", C.c_void_p), ("beta", C.c_void_p), ("mu_bandit", C.c_void_p),
        ("trans", C.c_double * 9),
        ("sigma_omega_initial", C.c_double), ("sigma_omega", C.c_double),
        ("change_omega_for_bandits", C.c_int32), ("_pad", C.c_int32),
        ("seed", C.c_uint64),
        ("time_mode", C.c_int32), ("_pad2", C.c_int32), ("time_mu", C.c_double), ("time_sigma", C.c_double),
    ]


class _Policy(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("table", C.c_void_p), ("epsilon", C.c_double)]


class _Log(C.Structure):
    _fields_ = [("cap", C.c_int64), ("n", C.c_int64),
                ("u", C.c_void_p), ("t", C.c_void_p), ("v", C.c_void_p), ("a", C.c_void_p),
                ("z", C.c_void_p), ("c", C.c_void_p), ("tf", C.c_void_p)]


class _Outputs(C.Structure):
    _fields_ = [("n_events", C.c_void_p), ("n_organic", C.c_void_p), ("n_bandit", C.c_void_p),
                ("n_clicks", C.c_void_p), ("last_view", C.c_void_p), ("hash", C.c_void_p),
                ("omega", C.c_void_p),
                ("view_hist", C.c_void_p), ("action_hist", C.c_void_p), ("click_hist", C.c_void_p),
                ("totals", C.c_void_p), ("log", C.POINTER(_Log))]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.rgo_exp.restype = C.c_double; L.rgo_exp.argtypes = [C.c_double]
        L.rgo_log.restype = C.c_double; L.rgo_log.argtypes = [C.c_double]
        L.rgo_norm_ppf.restype = C.c_double; L.rgo_norm_ppf.argtypes = [C.c_double]
        L.rgo_ff.restype = C.c_double; L.rgo_ff.argtypes = [C.c_double]
        L.rgo_philox.restype = None
        L.rgo_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.rgo_event_uniform.restype = C.c_double
        L.rgo_event_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]
        L.rgo_normals.restype = None
        L.rgo_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        L.rgo_policy_uniform_action.restype = C.c_int32
        L.rgo_policy_uniform_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32]
        L.rgo_time_normal.restype = C.c_double
        L.rgo_time_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.rgo_policy_uniform53.restype = C.c_double
        L.rgo_policy_uniform53.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]
        L.rgo_rollout.restype = None
        L.rgo_rollout.argtypes = [C.POINTER(_Env), C.POINTER(_Policy), C.c_int64, C.c_int64, C.c_int64,
                                  C.POINTER(_Outputs), C.c_int]
        L.rgo_batch_reset.restype = C.c_int64
        L.rgo_batch_reset.argtypes = [C.POINTER(_Env), C.c_int64] + [C.c_void_p] * 7 + [C.c_int64]
        L.rgo_batch_step.restype = C.c_int64
        L.rgo_batch_step.argtypes = [C.POINTER(_Env), C.c_int64] + [C.c_void_p] * 9 + [C.c_int64]
        L.rgo_view_probs.restype = None
        L.rgo_view_probs.argtypes = [C.POINTER(_Env), C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --- scalar helpers -------------------------------------------------------------------------
def philox(seed, user, t, slot):
    out = np.zeros(4, np.uint32)
    lib().rgo_philox(int(seed), int(user), int(t), int(slot), _p(out))
    return out


def exp(x): return lib().rgo_exp(float(x))
def log(x): return lib().rgo_log(float(x))
def norm_ppf(p): return lib().rgo_norm_ppf(float(p))
def ff(x): return lib().rgo_ff(float(x))
def event_uniform(seed, user, t, which): return lib().rgo_event_uniform(int(seed), int(user), int(t), int(which))
def policy_uniform_action(seed, user, t, P): return lib().rgo_policy_uniform_action(int(seed), int(user), int(t), int(P))


def time_normal(seed, user, call): return lib().rgo_time_normal(int(seed), int(user), int(call))


def policy_uniform53(seed, user, t, which): return lib().rgo_policy_uniform53(int(seed), int(user), int(t), int(which))


def normals(seed, user, t, K):
    z = np.zeros(K, np.float64)
    lib().rgo_normals(int(seed), int(user), int(t), int(K), _p(z))
    return z


# --- environment ----------------------------------------------------------------------------
class OracleEnv:
    """Static parameters of one reco-gym-v1 environment (host arrays, f64) + the seed."""

    def __init__(self, Gamma, mu_organic, beta, mu_bandit, trans, sigma_omega_initial, sigma_omega,
                 seed, change_omega_for_bandits=False, normal_time=None):
        """normal_time=(mu, sigma): NormalTimeGenerator clock (oracle only); None: the default event clock."""
        self.Gamma = np.ascontiguousarray(Gamma, np.float64)
        self.P, self.K = self.Gamma.shape
        self.mu_organic = np.ascontiguousarray(np.ravel(mu_organic), np.float64)
        self.beta = np.ascontiguousarray(beta, np.float64)
        self.mu_bandit = np.ascontiguousarray(np.ravel(mu_bandit), np.float64)
        self.trans = np.ascontiguousarray(trans, np.float64).reshape(9)
        assert self.beta.shape == (self.P, self.K) and self.mu_organic.shape == (self.P,)
        self.c = _Env()
        self.c.P, self.c.K = self.P, self.K
        self.c.Gamma, self.c.mu_organic = _p(self.Gamma), _p(self.mu_organic)
        self.c.beta, self.c.mu_bandit = _p(self.beta), _p(self.mu_bandit)
        for i in range(9):
            self.c.trans[i] = float(self.trans[i])
        self.c.sigma_omega_initial = float(sigma_omega_initial)
        self.c.sigma_omega = float(sigma_omega)
        self.c.change_omega_for_bandits = int(bool(change_omega_for_bandits))
        self.c.seed = int(seed)
        self.c.time_mode = 0 if normal_time is None else 1
        self.c.time_mu, self.c.time_sigma = (0.0, 1.0) if normal_time is None else (float(normal_time[0]), float(normal_time[1]))

    def view_probs(self, omega):
        omega = np.ascontiguousarray(omega, np.float64).ravel()
        out = np.zeros(self.P, np.float64)
        lib().rgo_view_probs(C.byref(self.c), _p(omega), _p(out))
        return out


def make_policy(kind="uniform", table=None, epsilon=0.0):
    pol = _Policy()
    if kind == "uniform":
        pol.kind = 0
        keep = None
    elif kind == "table":
        pol.kind = 1
        keep = np.ascontiguousarray(table, np.int32)
        pol.table = _p(keep)
    elif kind == "user_views":
        pol.kind = 2
        keep = None
    elif kind == "eps_table":       # EpsilonGreedy around a table agent (oracle only)
        pol.kind = 3
        pol.epsilon = float(epsilon)
        keep = np.ascontiguousarray(table, np.int32)
        pol.table = _p(keep)
    else:
        raise ValueError(kind)
    return pol, keep


def rollout(env, user_begin, n_users, policy="uniform", table=None, n_organic_users=0, log=False, log_cap=None,
            hist=False, omega=False, threads=1, epsilon=0.0):
    """generate_logs(n_users, agent, n_organic_users) on the oracle. Returns dict of numpy arrays."""
    total = n_users + n_organic_users
    pol, keep = make_policy(policy, table, epsilon)
    res = {
        "n_events": np.zeros(total, np.int32), "n_organic": np.zeros(total, np.int32),
        "n_bandit": np.zeros(total, np.int32), "n_clicks": np.zeros(total, np.int32),
        "last_view": np.zeros(total, np.int32), "hash": np.zeros(total, np.uint64),
        "totals": np.zeros(6, np.int64),
    }
    out = _Outputs()
    for k in ("n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "hash", "totals"):
        setattr(out, k, _p(res[k]))
    if omega:
        res["omega"] = np.zeros((total, env.K), np.float64)
        out.omega = _p(res["omega"])
    if hist:
        for k in ("view_hist", "action_hist", "click_hist"):
            res[k] = np.zeros(env.P, np.int64)
            setattr(out, k, _p(res[k]))
    lg = None
    if log:
        cap = int(log_cap or (total * 400 + 1000))
        lg = _Log()
        lg.cap, lg.n = cap, 0
        cols = {"u": np.zeros(cap, np.int32), "t": np.zeros(cap, np.int32), "v": np.zeros(cap, np.int32),
                "a": np.zeros(cap, np.int32), "z": np.zeros(cap, np.int8), "c": np.zeros(cap, np.int8),
                "tf": np.zeros(cap, np.float64)}
        for k, v in cols.items():
            setattr(lg, k, _p(v))
        out.log = C.pointer(lg)
    lib().rgo_rollout(C.byref(env.c), C.byref(pol), int(user_begin), int(n_organic_users), int(n_users),
                      C.byref(out), int(threads))
    if log:
        n = int(lg.n)
        if n > lg.cap:
            raise RuntimeError("oracle log capacity exceeded: %d > %d" % (n, lg.cap))
        res["log"] = {k: v[:n].copy() for k, v in cols.items()}
    del keep
    return res


def batch_reset(env, user_ids, sess_cap=None):
    user_ids = np.ascontiguousarray(user_ids, np.int64)
    N = user_ids.shape[0]
    cap = int(sess_cap or (N * 64 + 1024))
    st = {"user_id": user_ids, "state": np.zeros(N, np.int32), "t": np.zeros(N, np.int32),
          "last_view": np.zeros(N, np.int32), "omega": np.zeros((N, env.K), np.float64)}
    off = np.zeros(N + 1, np.int64)
    sv = np.zeros(cap, np.int32)
    n = lib().rgo_batch_reset(C.byref(env.c), N, _p(user_ids), _p(st["state"]), _p(st["t"]), _p(st["last_view"]),
                              _p(st["omega"]), _p(off), _p(sv), cap)
    assert n <= cap
    return st, off, sv[:n].copy()


def batch_step(env, st, action, sess_cap=None):
    N = st["user_id"].shape[0]
    action = np.ascontiguousarray(action, np.int32)
    cap = int(sess_cap or (N * 64 + 1024))
    reward = np.zeros(N, np.int32)
    off = np.zeros(N + 1, np.int64)
    sv = np.zeros(cap, np.int32)
    n = lib().rgo_batch_step(C.byref(env.c), N, _p(st["user_id"]), _p(action), _p(st["state"]), _p(st["t"]),
                             _p(st["last_view"]), _p(st["omega"]), _p(reward), _p(off), _p(sv), cap)
    assert n <= cap
    return reward, off, sv[:n].copy()


# --- groundwork for a device EpsilonGreedy policy ---------------------------------------------------------------------
def hole_cdf_table(P):
    """S[m] = the m-th partial sum of numpy's sequential cumsum over m copies of 1/(P-1) (binary64).  The cdf of
    EpsilonGreedy's exploring distribution (ones(P) with a 0 at the greedy action g, / sum; agents/epsilon_greedy.py:
    36-43) is then cdf[j] = S[j+1] for j < g and S[j] for j >= g: adding the hole's exact 0 changes nothing, so ONE table
    per P serves every g, and choice(P, p) becomes a binary search in S instead of P sequential additions per draw."""
    c = 1.0 / float(P - 1)
    S = np.zeros(P, np.float64)
    acc = 0.0
    for m in range(1, P):
        acc = acc + c
        S[m] = acc
    return S


def hole_choice(S, g, u):
    """numpy legacy choice(P, p=uniform-with-hole-at-g) on the uniform u, from the table of hole_cdf_table."""
    P = len(S)
    last = S[P - 1]
    lo, hi = 0, P                       # smallest j with cdf[j] / last > u
    while lo < hi:
        mid = (lo + hi) // 2
        cdf = S[mid + 1] if mid < g else S[mid]
        if cdf / last <= u:
            lo = mid + 1
        else:
            hi = mid
    return lo

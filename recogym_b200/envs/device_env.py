"""DeviceEnv: the static parameters of one reco-gym-v1 environment resident on a B200, plus the
vectorised entry points (fused rollout, batched reset/step) that call the CUDA kernels through the
C ABI (include/recogym_b200.h).  torch supplies device memory and streams only.
"""
import ctypes as C

import numpy as np

from .. import _lib as L


def _torch():
    import torch
    return torch


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceEnv:
    """Owns an `rg_env` handle (device copies of Gamma, mu, beta and the transition thresholds)."""

    def __init__(self, static_params, sigma_omega_initial, sigma_omega, seed, change_omega_for_bandits=False,
                 device=0, normal_time=None):
        """normal_time: None = DefaultTimeGenerator (the clock is the event index); (mu, sigma) = NormalTimeGenerator
        (normal_time_generator.py:7-31): the clock advances by |N(mu, sigma)| per event (Philox slot 2 at (user, event
        index)), the omega drift is scaled by that step (reco_env_v1.py:89-99)."""
        self.Gamma = np.ascontiguousarray(static_params["Gamma"], np.float64)
        self.P, self.K = self.Gamma.shape
        self.mu_organic = np.ascontiguousarray(np.ravel(static_params["mu_organic"]), np.float64)
        self.beta = np.ascontiguousarray(static_params["beta"], np.float64)
        self.mu_bandit = np.ascontiguousarray(np.ravel(static_params["mu_bandit"]), np.float64)
        self.state_transition = np.ascontiguousarray(static_params["state_transition"], np.float64)
        self.device = int(device)
        self.seed = int(seed)
        d = L.EnvDesc()
        d.num_products, d.K = self.P, self.K
        d.Gamma = self.Gamma.ctypes.data
        d.mu_organic = self.mu_organic.ctypes.data
        d.beta = self.beta.ctypes.data
        d.mu_bandit = self.mu_bandit.ctypes.data
        for i, v in enumerate(self.state_transition.reshape(9)):
            d.state_transition[i] = float(v)
        d.sigma_omega_initial = float(sigma_omega_initial)
        d.sigma_omega = float(sigma_omega)
        d.change_omega_for_bandits = int(bool(change_omega_for_bandits))
        d.seed = self.seed
        self.normal_time = None if normal_time is None else (float(normal_time[0]), float(normal_time[1]))
        if self.normal_time is not None:
            d.time_generator, d.normal_time_mu, d.normal_time_sigma = 1, self.normal_time[0], self.normal_time[1]
        h = C.c_void_p()
        L.check(L.lib().rg_env_create(C.byref(d), self.device, C.byref(h)))
        self._h = h
        self._desc_args = (sigma_omega_initial, sigma_omega, change_omega_for_bandits)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.lib().rg_env_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clone(self):
        """Independent handle with the same parameters (deepcopy support of the facade)."""
        sp = {"Gamma": self.Gamma, "mu_organic": self.mu_organic, "beta": self.beta, "mu_bandit": self.mu_bandit,
              "state_transition": self.state_transition}
        return DeviceEnv(sp, self._desc_args[0], self._desc_args[1], self.seed, self._desc_args[2], self.device,
                         self.normal_time)

    def set_seed(self, seed):
        """AbstractEnv.reset_random_seed(epoch): seed = random_seed + epoch (abstract.py:59-62)."""
        self.seed = int(seed)
        L.check(L.lib().rg_env_set_seed(self._h, self.seed))

    @property
    def torch_device(self):
        return _torch().device("cuda", self.device)

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.torch_device).cuda_stream)

    # -- fused rollout -------------------------------------------------------------------------
    def rollout_device(self, n_users, policy="uniform", table=None, user_begin=0, n_organic_users=0,
                       math_mode=L.MATH_FAST, per_user=True, omega=False, hist=False, totals=None, epsilon=0.0):
        """generate_logs(n_users, agent, n_organic_users) reduced to counters, everything on device.
        policy: "uniform" | "table" | "user_views" | "eps_table" (EpsilonGreedy(epsilon) around the table agent).

        Returns a dict of torch tensors (device): per-user n_events/n_organic/n_bandit/n_clicks/
        last_view/hash (if per_user), omega (if requested), view/action/click histograms (if hist)
        and totals[8] (int64, accumulated into `totals` when given).  Asynchronous on the current stream.
        """
        torch = _torch()
        dev = self.torch_device
        n = int(n_users) + int(n_organic_users)
        res = {}
        out = L.RolloutOut()
        if per_user:
            for name in ("n_events", "n_organic", "n_bandit", "n_clicks", "last_view"):
                res[name] = torch.empty(n, dtype=torch.int32, device=dev)
                setattr(out, name, res[name].data_ptr())
            res["hash"] = torch.empty(n, dtype=torch.int64, device=dev)
            out.hash = res["hash"].data_ptr()
        if omega:
            res["omega"] = torch.empty((n, self.K), dtype=torch.float64, device=dev)
            out.omega = res["omega"].data_ptr()
        if hist:
            for name in ("view_hist", "action_hist", "click_hist"):
                res[name] = torch.zeros(self.P, dtype=torch.int64, device=dev)
                setattr(out, name, res[name].data_ptr())
        res["totals"] = totals if totals is not None else torch.zeros(16, dtype=torch.int64, device=dev)
        out.totals = res["totals"].data_ptr()
        if policy in ("table", "eps_table") and not (hasattr(table, "is_cuda") and table.is_cuda):
            table = torch.as_tensor(np.ascontiguousarray(table, np.int32), device=dev)
        keep = res["_keep"] = []  # keeps the table alive until the launch has been enqueued
        pol = _policy_struct(keep, policy, table, self.P, host=False, epsilon=epsilon)
        L.check(L.lib().rg_rollout_device(self._h, C.byref(pol), int(user_begin), int(n_organic_users),
                                          int(n_users), C.byref(out), int(math_mode), self._stream()))
        return res

    def rollout_host(self, n_users, policy="uniform", table=None, user_begin=0, n_organic_users=0,
                     math_mode=L.MATH_FAST, per_user=True, omega=False, hist=False, epsilon=0.0):
        """Same through HOST (numpy) buffers: the C ABI performs the H2D/D2H copies. Synchronous."""
        n = int(n_users) + int(n_organic_users)
        res = {}
        out = L.RolloutOut()
        if per_user:
            for name in ("n_events", "n_organic", "n_bandit", "n_clicks", "last_view"):
                res[name] = np.zeros(n, np.int32)
                setattr(out, name, res[name].ctypes.data)
            res["hash"] = np.zeros(n, np.uint64)
            out.hash = res["hash"].ctypes.data
        if omega:
            res["omega"] = np.zeros((n, self.K), np.float64)
            out.omega = res["omega"].ctypes.data
        if hist:
            for name in ("view_hist", "action_hist", "click_hist"):
                res[name] = np.zeros(self.P, np.int64)
                setattr(out, name, res[name].ctypes.data)
        res["totals"] = np.zeros(16, np.int64)
        out.totals = res["totals"].ctypes.data
        keep = []
        pol = _policy_struct(keep, policy, table, self.P, host=True, epsilon=epsilon)
        L.check(L.lib().rg_rollout_host(self._h, C.byref(pol), int(user_begin), int(n_organic_users), int(n_users),
                                        C.byref(out), int(math_mode)))
        del keep
        return res


def _policy_struct(keep, policy, table, P, host, epsilon=0.0):
    pol = L.Policy()
    pol.epsilon = float(epsilon)
    if policy == "uniform":
        pol.kind = L.POLICY_UNIFORM
    elif policy == "user_views":   # OrganicUserEventCounterAgent, default configuration (RG_POLICY_USER_VIEWS)
        pol.kind = L.POLICY_USER_VIEWS
    elif policy in ("table", "eps_table"):   # eps_table: EpsilonGreedy(epsilon) around the table agent (RG_POLICY_EPS_TABLE)
        pol.kind = L.POLICY_TABLE if policy == "table" else L.POLICY_EPS_TABLE
        if host:
            if hasattr(table, "is_cuda"):     # a table trained on the device (CountStats): the host path takes a host copy
                table = table.detach().cpu().numpy()
            t = np.ascontiguousarray(table, np.int32)
            assert t.shape == (P,)
            pol.table = t.ctypes.data
        else:
            torch = _torch()
            t = table if (hasattr(table, "is_cuda") and table.is_cuda) else None
            assert t is not None and t.dtype == torch.int32 and t.numel() == P
            pol.table = t.data_ptr()
        keep.append(t)
    else:
        raise ValueError("unknown policy %r" % (policy,))
    return pol


def rollout_log(env, n_users, policy="uniform", table=None, user_begin=0, n_organic_users=0,
                math_mode=L.MATH_TENSOR, rows_per_user=110.0, epsilon=0.0):
    """The rows of AbstractEnv.generate_logs (abstract.py:241-327) for users user_begin.., produced by ONE fused
    rollout (event-log output of the kernels) and sorted into the reference order on the host.
    Returns a dict of numpy arrays: u, t, z (0 organic / 1 bandit), v, a, c (-1 where the reference has NaN)."""
    n = int(n_users) + int(n_organic_users)
    cap = int(n * rows_per_user) + 4096
    keep = []
    while True:
        log = np.empty((cap, 4), np.int32)
        count = np.zeros(1, np.int64)
        totals = np.zeros(16, np.int64)
        out = L.RolloutOut()
        out.totals = totals.ctypes.data
        out.log, out.log_capacity, out.log_count = log.ctypes.data, cap, count.ctypes.data
        pol = _policy_struct(keep, policy, table, env.P, host=True, epsilon=epsilon)
        L.check(L.lib().rg_rollout_host(env._h, C.byref(pol), int(user_begin), int(n_organic_users), int(n_users),
                                        C.byref(out), int(math_mode)))
        rows = int(count[0])
        if rows <= cap:
            break
        cap = rows + 1024  # the buffer was too small: the rows are a pure function of the inputs, run again
    log = log[:rows]
    order = np.lexsort((log[:, 1], log[:, 0]))
    log = log[order]
    z = (log[:, 3] & 1).astype(np.int8)
    click = ((log[:, 3] >> 1) & 1).astype(np.int8)
    return {"u": log[:, 0] + int(user_begin), "t": log[:, 1], "z": z,
            "v": np.where(z == 0, log[:, 2], -1).astype(np.int32), "a": np.where(z == 1, log[:, 2], -1).astype(np.int32),
            "c": np.where(z == 1, click, -1).astype(np.int8), "totals": totals}


class BatchState:
    """Device SoA of N concurrent users for the batched gym protocol."""

    def __init__(self, env, user_ids, sess_cap=32):
        torch = _torch()
        dev = env.torch_device
        self.env = env
        self.user_id = torch.as_tensor(np.ascontiguousarray(user_ids, np.int64), device=dev)
        self.N = int(self.user_id.numel())
        self.sess_cap = int(sess_cap)
        # one int32 slab: state | t | last_view | reward | sess_len | sess_v   (a single D2H reads it all)
        self.slab = torch.zeros(self.N * (5 + self.sess_cap), dtype=torch.int32, device=dev)
        N = self.N
        self.state = self.slab[0:N]
        self.t = self.slab[N:2 * N]
        self.last_view = self.slab[2 * N:3 * N]
        self.reward = self.slab[3 * N:4 * N]
        self.sess_len = self.slab[4 * N:5 * N]
        self.sess_v = self.slab[5 * N:].view(N, self.sess_cap)
        self.omega = torch.zeros((N, env.K), dtype=torch.float64, device=dev)
        self.n_unfinished = torch.zeros(1, dtype=torch.int32, device=dev)

    def _common(self):
        return (_ptr(self.state), _ptr(self.t), _ptr(self.last_view), _ptr(self.omega))

    def reset(self, math_mode=L.MATH_EXACT):
        """env.reset(user) + env.step(None) for every user (abstract.py:90-103,157-172)."""
        e = self.env
        self.n_unfinished.zero_()
        L.check(L.lib().rg_batch_reset_device(e._h, self.N, _ptr(self.user_id), *self._common(), _ptr(self.sess_v),
                                              _ptr(self.sess_len), self.sess_cap, _ptr(self.n_unfinished),
                                              int(math_mode), e._stream()))

    def step(self, action, math_mode=L.MATH_EXACT):
        """env.step(action) for every user not yet done (abstract.py:174-197)."""
        torch = _torch()
        e = self.env
        if not (hasattr(action, "is_cuda") and action.is_cuda):
            action = torch.as_tensor(np.ascontiguousarray(action, np.int32), device=e.torch_device)
        assert action.dtype == torch.int32 and action.numel() == self.N
        self.n_unfinished.zero_()
        L.check(L.lib().rg_batch_step_device(e._h, self.N, _ptr(self.user_id), _ptr(action), *self._common(),
                                             _ptr(self.reward), _ptr(self.sess_v), _ptr(self.sess_len),
                                             self.sess_cap, _ptr(self.n_unfinished), int(math_mode), e._stream()))

    def continue_sessions(self, math_mode=L.MATH_EXACT):
        """more views for users whose organic session did not fit in sess_cap (abstract.py:105-121)."""
        e = self.env
        self.n_unfinished.zero_()
        L.check(L.lib().rg_batch_continue_device(e._h, self.N, _ptr(self.user_id), *self._common(),
                                                 _ptr(self.sess_v), _ptr(self.sess_len), self.sess_cap,
                                                 _ptr(self.n_unfinished), int(math_mode), e._stream()))

    def read(self):
        """One D2H copy of the int32 slab -> dict of numpy views (synchronises)."""
        h = self.slab.cpu().numpy()
        N = self.N
        return {"state": h[0:N], "t": h[N:2 * N], "last_view": h[2 * N:3 * N], "reward": h[3 * N:4 * N],
                "sess_len": h[4 * N:5 * N], "sess_v": h[5 * N:].reshape(N, self.sess_cap)}

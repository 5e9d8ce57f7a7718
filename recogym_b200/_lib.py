"""ctypes binding of recogym_b200/librecogym_b200.so (the C ABI of include/recogym_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C recogym_b200/csrc`.
There is no CPU fallback anywhere: a missing library raises at import of the first symbol, a
missing GPU raises RecoGymB200Error(RG_ENODEVICE) from rg_env_create.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RG_LIB_PATH") or os.path.join(_HERE, "librecogym_b200.so")   # override: kernel experiments

RG_OK, RG_EINVAL, RG_ECUDA, RG_ENOMEM, RG_ENODEVICE = 0, -1, -2, -3, -4
ORGANIC, BANDIT, STOP = 0, 1, 2
POLICY_UNIFORM, POLICY_TABLE, POLICY_USER_VIEWS, POLICY_EPS_TABLE = 0, 1, 2, 3
MATH_EXACT, MATH_FAST, MATH_TENSOR = 0, 1, 2
TRAIN_TEST_AGENT, TRAIN_EVOLUTION_STEP, TRAIN_EVOLUTION_INIT = 0, 1, 2


class RecoGymB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("recogym_b200 error %d: %s" % (code, text))
        self.code = code


class EnvDesc(C.Structure):
    _fields_ = [
        ("num_products", C.c_int32), ("K", C.c_int32),
        ("Gamma", C.c_void_p), ("mu_organic", C.c_void_p), ("beta", C.c_void_p), ("mu_bandit", C.c_void_p),
        ("state_transition", C.c_double * 9),
        ("sigma_omega_initial", C.c_double), ("sigma_omega", C.c_double),
        ("change_omega_for_bandits", C.c_int32), ("reserved", C.c_int32),
        ("seed", C.c_uint64),
        ("time_generator", C.c_int32), ("reserved2", C.c_int32),
        ("normal_time_mu", C.c_double), ("normal_time_sigma", C.c_double),
    ]


class Policy(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("table", C.c_void_p), ("epsilon", C.c_double)]


class RolloutOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "hash", "omega",
        "view_hist", "action_hist", "click_hist", "totals")] + [
        ("log", C.c_void_p), ("log_capacity", C.c_int64), ("log_count", C.c_void_p)]


class LogColumns(C.Structure):
    _fields_ = [("capacity", C.c_int64)] + [(n, C.c_void_p) for n in (
        "t", "u", "z", "v", "a", "c", "ps", "lv", "lvp", "sess", "flags")]


class Env0Desc(C.Structure):       # rg_env0_desc (reco-gym-v0)
    _fields_ = [("num_products", C.c_int32), ("num_clusters", C.c_int32), ("init_cdf", C.c_void_p), ("view_cdf", C.c_void_p),
                ("cluster_of", C.c_void_p), ("click_thr", C.c_void_p), ("trans_cdf", C.c_double * 9), ("seed", C.c_uint64),
                ("device", C.c_int32), ("reserved", C.c_int32)]


class Env0Out(C.Structure):        # rg_env0_out
    _fields_ = [(n, C.c_void_p) for n in ("totals", "n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "state",
                                          "log")] + [("log_capacity", C.c_int64), ("log_count", C.c_void_p)]


POLICY0_UNIFORM, POLICY0_TABLE, POLICY0_REPLAY = 0, 1, 2


class SparseCounts(C.Structure):
    _fields_ = [("capacity", C.c_int64)] + [(n, C.c_void_p) for n in ("keys", "a", "b", "row_a", "row_b", "stats")]


_lib = None

EXPORTS = (
    "rg_env_create", "rg_env_set_seed", "rg_env_destroy", "rg_rollout_device", "rg_rollout_host",
    "rg_generate_logs_device", "rg_count_stats_device", "rg_count_policy_device", "rg_ips_device",
    "rg_recall_device",
    "rg_batch_reset_device", "rg_batch_step_device", "rg_batch_continue_device",
    "rg_spec_philox", "rg_spec_exp", "rg_spec_log", "rg_spec_norm_ppf", "rg_spec_ff",
    "rg_spec_event_uniform", "rg_spec_normals", "rg_spec_policy_uniform_action", "rg_spec_time_step",
    "rg_spec_eval_device", "rg_tc_logits_device", "rg_device_count", "rg_version", "rg_last_error",
    "rg_debug_force_tier4", "rg_sparse_count_stats_device", "rg_sparse_count_policy_device",
    "rg_sparse_rehash_device",
    "rg_env0_create", "rg_env0_destroy", "rg_env0_set_seed", "rg_env0_rollout_device", "rg_env0_rollout_host",
    "rg_spec_env0_rollout",
)


def lib():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "recogym_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C recogym_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    L.rg_last_error.restype = C.c_char_p
    L.rg_version.restype = C.c_int
    L.rg_device_count.restype = C.c_int
    L.rg_env_create.restype = C.c_int
    L.rg_env_create.argtypes = [C.POINTER(EnvDesc), C.c_int, C.POINTER(vp)]
    L.rg_env_set_seed.restype = C.c_int
    L.rg_env_set_seed.argtypes = [vp, u64]
    L.rg_env_destroy.restype = C.c_int
    L.rg_env_destroy.argtypes = [vp]
    L.rg_rollout_device.restype = C.c_int
    L.rg_rollout_device.argtypes = [vp, C.POINTER(Policy), i64, i64, i64, C.POINTER(RolloutOut), C.c_int, vp]
    L.rg_rollout_host.restype = C.c_int
    L.rg_rollout_host.argtypes = [vp, C.POINTER(Policy), i64, i64, i64, C.POINTER(RolloutOut), C.c_int]
    L.rg_generate_logs_device.restype = C.c_int
    L.rg_generate_logs_device.argtypes = [vp, C.POINTER(Policy), vp, i64, i64, i64, i32, C.c_int, vp,
                                          C.POINTER(LogColumns), vp, vp, vp, vp, C.c_int, vp]
    L.rg_count_stats_device.restype = C.c_int
    L.rg_count_stats_device.argtypes = [C.POINTER(LogColumns), i64, i32, C.c_int, vp, vp, vp, vp]
    L.rg_count_policy_device.restype = C.c_int
    L.rg_count_policy_device.argtypes = [C.c_int, i32, vp, vp, vp, vp, vp]
    L.rg_sparse_count_stats_device.restype = C.c_int
    L.rg_sparse_count_stats_device.argtypes = [C.POINTER(LogColumns), i64, C.c_int, C.POINTER(SparseCounts),
                                               C.POINTER(SparseCounts), vp]
    L.rg_sparse_rehash_device.restype = C.c_int
    L.rg_sparse_rehash_device.argtypes = [C.POINTER(SparseCounts), C.POINTER(SparseCounts), vp]
    L.rg_sparse_count_policy_device.restype = C.c_int
    L.rg_sparse_count_policy_device.argtypes = [C.c_int, i32, C.POINTER(SparseCounts), vp, vp, vp]
    L.rg_ips_device.restype = C.c_int
    L.rg_ips_device.argtypes = [C.POINTER(LogColumns), i64, i64, vp, vp, vp]
    L.rg_recall_device.restype = C.c_int
    L.rg_recall_device.argtypes = [C.POINTER(LogColumns), i64, i64, vp, vp, i32, vp, vp]
    L.rg_batch_reset_device.restype = C.c_int
    L.rg_batch_reset_device.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, i32, vp, C.c_int, vp]
    L.rg_batch_step_device.restype = C.c_int
    L.rg_batch_step_device.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, C.c_int, vp]
    L.rg_batch_continue_device.restype = C.c_int
    L.rg_batch_continue_device.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, i32, vp, C.c_int, vp]
    L.rg_spec_philox.restype = None
    L.rg_spec_philox.argtypes = [u64, u64, u32, u32, vp]
    for name in ("rg_spec_exp", "rg_spec_log", "rg_spec_norm_ppf", "rg_spec_ff"):
        getattr(L, name).restype = dbl
        getattr(L, name).argtypes = [dbl]
    L.rg_spec_event_uniform.restype = dbl
    L.rg_spec_event_uniform.argtypes = [u64, u64, u32, C.c_int]
    L.rg_spec_normals.restype = None
    L.rg_spec_normals.argtypes = [u64, u64, u32, C.c_int, vp]
    L.rg_spec_policy_uniform_action.restype = i32
    L.rg_spec_policy_uniform_action.argtypes = [u64, u64, u32, i32]
    L.rg_spec_time_step.restype = dbl
    L.rg_spec_time_step.argtypes = [u64, u64, u32, dbl, dbl]
    L.rg_spec_eval_device.restype = C.c_int
    L.rg_spec_eval_device.argtypes = [C.c_int, i64, vp, vp, vp]
    L.rg_debug_force_tier4.restype = C.c_int
    L.rg_debug_force_tier4.argtypes = [C.c_int]
    L.rg_tc_logits_device.restype = C.c_int
    L.rg_tc_logits_device.argtypes = [vp, vp, vp, vp, vp]
    L.rg_env0_create.restype = C.c_int
    L.rg_env0_create.argtypes = [C.POINTER(Env0Desc), C.POINTER(vp)]
    L.rg_env0_destroy.restype = None
    L.rg_env0_destroy.argtypes = [vp]
    L.rg_env0_set_seed.restype = C.c_int
    L.rg_env0_set_seed.argtypes = [vp, u64]
    L.rg_env0_rollout_device.restype = C.c_int
    L.rg_env0_rollout_device.argtypes = [vp, C.c_int, vp, i64, i64, i64, i64, C.POINTER(Env0Out), vp]
    L.rg_env0_rollout_host.restype = C.c_int
    L.rg_env0_rollout_host.argtypes = [vp, C.c_int, vp, i64, i64, i64, i64, C.POINTER(Env0Out)]
    L.rg_spec_env0_rollout.restype = C.c_int
    L.rg_spec_env0_rollout.argtypes = [C.POINTER(Env0Desc), C.c_int, vp, i64, i64, i64, i64, C.POINTER(Env0Out)]
    _lib = L
    return L


def check(rc):
    if rc != RG_OK:
        raise RecoGymB200Error(rc, lib().rg_last_error().decode("utf-8", "replace"))


def device_count():
    return int(lib().rg_device_count())

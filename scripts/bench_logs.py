#!/usr/bin/env python
"""Micro-benchmark of the log pipeline either side of the rollout (csrc/rg_logs.cu), HBM-bound integer kernels:
prints one JSON line per stage with the achieved GB/s over ALGORITHMIC bytes against MEASURED_PEAKS.json hbm_gbs.
usage: python scripts/bench_logs.py [--users N] [--P P] [--K K]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--P", type=int, default=100)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--df-users", type=int, default=100_000, help="users of the generate_logs -> pandas DataFrame end-to-end line")
    a = ap.parse_args()
    import torch
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.envs.static_params import draw_static_params
    from recogym_b200.logs import CountStats, generate_logs_device, ips_table_policy
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    cfg = rg.Configuration({**rg.env_1_args, "num_products": a.P, "K": a.K, "random_seed": 42})
    sp = draw_static_params(cfg)
    de = rg.DeviceEnv(sp, 1, 0.1, 42)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=a.reps):
        fn()
        torch.cuda.synchronize()
        ms = []
        for r in range(reps):
            flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    def line(stage, ms, bytes_alg, extra=None):
        gbs = bytes_alg / (ms * 1e-3) / 1e9
        print(json.dumps({"stage": stage, "ms": ms, "algorithmic_bytes": bytes_alg, "achieved_GBps": gbs,
                          "peak_GBps": peak, "frac": gbs / peak, **(extra or {})}))

    log = generate_logs_device(de, a.users, "uniform", None, None, 0, 0, L.MATH_TENSOR)
    rows = log.rows
    ms_roll = timed(lambda: de.rollout_device(a.users, "uniform", None, 0, 0, L.MATH_TENSOR, per_user=False))
    ms_logs = timed(lambda: generate_logs_device(de, a.users, "uniform", None, None, 0, 0, L.MATH_TENSOR, log=log))
    # order + columns on top of the rollout: 16 B record written by the rollout and read by the scatter, 16 B written +
    # read ordered, 41 B of columns written per row, 4+8 B per user for the scan
    extra_bytes = rows * (16 + 16 + 16 + 16 + 41) + a.users * 20
    line("rollout_only", ms_roll, 0, {"users": a.users, "rows": rows, "user_steps_per_s": (rows - a.users) / (ms_roll * 1e-3)})
    line("generate_logs_device(rollout+order+columns)", ms_logs, extra_bytes,
         {"rows_per_s": rows / (ms_logs * 1e-3), "ms_over_rollout": ms_logs - ms_roll,
          "order_columns_GBps": extra_bytes / max(ms_logs - ms_roll, 1e-6) / 1e6})
    st = CountStats(de.P)
    n_org_rows = int((log.z[:rows] == 0).sum())
    ms_st = timed(lambda: st.train(log), reps=3)
    # reads z,v,sess (9 B) per row + z,a,c,lvp (13 B) per row; atomics: ~6.7 pairs x 2 per organic row + 1-2 per bandit row
    line("count_stats(train)", ms_st, rows * 22, {"organic_rows": n_org_rows, "rows_per_s": rows / (ms_st * 1e-3)})
    ms_am = timed(lambda: (st.organic_count_policy(), st.bandit_count_policy()))
    line("count_policy(argmax x2)", ms_am, 3 * 4 * a.P * a.P)
    table = st.organic_count_policy()["table"]
    ms_ips = timed(lambda: ips_table_policy(log, table))
    line("ips_table_policy", ms_ips, rows * 25)
    # end to end through the reference-facing call: env.generate_logs(n) -> pandas DataFrame (fused rollout + ordered
    # device log + device->host copy of the columns + DataFrame construction), wall clock
    import time
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "num_products": a.P, "K": a.K, "random_seed": 42})
    env.generate_logs(1000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    df = env.generate_logs(a.df_users)
    dt = time.perf_counter() - t0
    print(json.dumps({"stage": "env.generate_logs -> pandas DataFrame (end to end, wall clock)", "users": a.df_users,
                      "rows": int(len(df)), "seconds": dt, "rows_per_s": len(df) / dt,
                      "user_steps_per_s": (len(df) - a.df_users) / dt,
                      "d2h_bytes": int(df.memory_usage(index=False, deep=False).sum())}))


if __name__ == "__main__":
    main()

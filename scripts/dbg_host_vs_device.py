#!/usr/bin/env python
"""Per-call wall time of rg_rollout_host against rg_rollout_device (+ synchronize) on the same users.
usage: python scripts/dbg_host_vs_device.py [--P 100000 --K 64 --users 200000]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=100000); ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--users", type=int, default=200000)
    a = ap.parse_args()
    import torch
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.envs.static_params import draw_static_params
    cfg = rg.Configuration({**rg.env_1_args, "num_products": a.P, "K": a.K, "random_seed": 42})
    sp = draw_static_params(cfg)
    de = rg.DeviceEnv(sp, cfg.sigma_omega_initial, cfg.sigma_omega, 42)
    tab = np.random.RandomState(0).randint(0, a.P, a.P).astype(np.int32)
    tdev = torch.as_tensor(tab, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = de.rollout_device(a.users, "table", tdev, rep * a.users, 0, L.MATH_TENSOR, per_user=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        h = de.rollout_host(a.users, "table", tab, rep * a.users, 0, L.MATH_TENSOR, per_user=False)
        t2 = time.perf_counter()
        assert np.array_equal(r["totals"].cpu().numpy()[:6], h["totals"][:6])
        print("rep %d: device %.1f ms, host %.1f ms, events %d" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), int(h["totals"][1])), flush=True)

if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Pass-1 hand-shake profile of the tensor-mode rollout (experiment build, `make -C recogym_b200/csrc prof`):

    RG_LIB_PATH=recogym_b200/librecogym_b200_prof.so python scripts/prof_pass1.py [--P 10000 --K 20 --users 1000000]

Prints the cycles every pass-1 role (MMA issuer, TMA loader, epilogue warps) spends waiting for the others, per
128 x 128 accumulator tile, and the time line of round 30 of CTA 0.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=10000)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--tag", default="")
    ap.add_argument("--trained", action="store_true", help="act with bench.py's trained organic_count table instead of a random one")
    ap.add_argument("--flush", action="store_true", help="flush L2 (256 MiB write) before the timed launch, as bench.py does")
    ap.add_argument("--user-begin", type=int, default=0)
    args = ap.parse_args()
    import torch
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.envs.static_params import draw_static_params
    cfg = rg.Configuration({**rg.env_1_args, "num_products": args.P, "K": args.K, "random_seed": 42})
    sp = draw_static_params(cfg)
    de = rg.DeviceEnv(sp, cfg.sigma_omega_initial, cfg.sigma_omega, 42)
    if args.trained:
        import bench
        pol = bench.build_policies(["organic_count"], args.P, args.K, 1000, device=0)
        table = torch.as_tensor(pol[0][2], device="cuda")
    else:
        table = torch.as_tensor(np.random.RandomState(0).randint(0, args.P, args.P).astype(np.int32), device="cuda")
    lib = L.lib()
    has_prof = hasattr(lib, "rg_prof_read")
    de.rollout_device(args.users // 4, "table", table, 5_000_000, 0, L.MATH_TENSOR, per_user=False)
    torch.cuda.synchronize()
    prof = (C.c_uint64 * 64)()
    trace = (C.c_int64 * 8192)()
    if has_prof:
        lib.rg_prof_read(prof, None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.flush:
        fl = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        fl.fill_(1)
    e0.record()
    r = de.rollout_device(args.users, "table", table, args.user_begin, 0, L.MATH_TENSOR, per_user=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tot = r["totals"].cpu().numpy()
    out = {"tag": args.tag, "P": args.P, "K": args.K, "users": args.users, "ms": ms, "user_steps_per_s": float(tot[1]) / ms * 1e3,
           "organic_draws": int(tot[2]),
           "phase_share": dict(zip(("compact", "pass1", "pass2", "resample", "advance", "omega"),
                                   [round(float(x) / max(1.0, float(tot[10:16].sum())), 4) for x in tot[10:16]])),
           "pass1_cycles_per_128_draws": float(tot[11]) / max(1.0, float(tot[2])) * 128.0}
    if has_prof:
        lib.rg_prof_read(prof, trace, 0)
        p = np.array(list(prof), dtype=np.float64)
        ep, iss, ld = p[0:16], p[16:32], p[32:48]
        ntile_i = max(iss[3], 1.0)
        ntile_e = max(ep[9], 1.0)
        out["issuer_per_unit"] = {"wait_tma": iss[0] / ntile_i, "wait_dempty": iss[1] / ntile_i, "issue": iss[2] / ntile_i,
                                  "units": iss[3]}
        out["loader_per_unit"] = {"wait_sfree_plus_issue": ld[4] / max(ld[5], 1.0), "units": ld[5]}
        out["epilogue_per_tile_per_warp"] = {"wait_dfull": ep[6] / ntile_e, "tmem_ld": ep[7] / ntile_e,
                                             "tile_total_after_dfull": ep[8] / ntile_e, "tiles_x_warps": ep[9]}
        out["pass1_thread0_per_round"] = {"a_tile_build_and_sync": ep[10] / max(ep[13], 1.0), "tile_loop": ep[11] / max(ep[13], 1.0),
                                          "tail_sync_and_omega_tile": ep[14] / max(ep[13], 1.0)}
        sub = p[48:58]
        names = ("A_compact", "A_bounds", "B_pass1", "C_pass2a", "C_pass2b", "C2_resample", "D1_view", "D2_bandit_runs",
                 "O_list", "O_normals")
        out["thread0_cycles_per_round"] = {n: round(float(v) / max(ep[13], 1.0), 1) for n, v in zip(names, sub)}
        out["rounds"] = ep[13]
        out["mean_live_per_round"] = ep[12] / max(ep[13], 1.0)
        t = np.array(list(trace), dtype=np.int64).reshape(-1, 8)
        nz = t[:, 2] != 0
        if nz.any():
            t0 = t[nz][:, [0, 1, 2]].min()
            rows = []
            for i in np.nonzero(nz)[0][:40]:
                rows.append([int(i)] + [int(x - t0) if x else None for x in t[i]])
            out["trace_cols"] = ["tile", "iss_tma_ok", "iss_dempty_ok", "iss_issued", "epi0_dfull_ok", "epi0_done",
                                 "epi3_dfull_ok", "loader_issued", "epi3_done"]
            out["trace"] = rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()

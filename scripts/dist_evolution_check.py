#!/usr/bin/env python
"""Multi-GPU evolution loop == single-GPU evolution loop (training statistics reduced at step
boundaries).  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port 29541 scripts/dist_evolution_check.py.  Every rank runs evaluate_agent twice -- sharded over the ranks and
alone (sharded=False: all users on this GPU) -- for the count agents, every TrainingApproach, dense and sparse statistics,
and compares the reward dicts and the final acting tables; rank 0 prints one JSON line."""
import copy
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import torch.distributed as dist
    import recogym_b200 as rg
    import importlib
    ea = importlib.import_module("recogym_b200.evaluate_agent")   # (the package exports the FUNCTION under this name)
    from recogym_b200 import logs
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = []
    T = rg.TrainingApproach

    def run(tag, P, K, make_agent, approach, n_init, n_step, n_steps, window=10000, sparse_min=None):
        old = logs.SPARSE_MIN_PRODUCTS
        if sparse_min is not None:
            logs.SPARSE_MIN_PRODUCTS = sparse_min
        try:
            out = []
            dt = []
            for sharded in (True, False):
                env = rg.make("reco-gym-v1")
                env.init_gym({**rg.env_1_args, "random_seed": 11, "num_products": P, "K": K, "device": local})
                agent = make_agent(P)
                inner = ea._device_evolution_agent(agent)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = ea._evaluate_agent_device(env, agent, inner, n_init, n_step, n_steps, approach,
                                              sliding_window_samples=window, sharded=sharded)
                torch.cuda.synchronize()
                dt.append(time.perf_counter() - t0)
                # the loop hands the trained copy back as the last `training` deepcopy's source: the acting table of a
                # further step is what matters
                out.append(r)
            a, b = out
            same = all(a[k] == b[k] for k in a)
            cases.append({"case": tag, "equal": bool(same), "successes": a[rg.EvolutionCase.SUCCESS],
                          "failures": a[rg.EvolutionCase.FAILURE], "s_sharded": round(dt[0], 3), "s_single": round(dt[1], 3)})
            if rank == 0:
                print("[dist_evolution_check]", cases[-1], file=sys.stderr, flush=True)
        finally:
            logs.SPARSE_MIN_PRODUCTS = old

    oc = lambda P: rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    bc = lambda P: rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    eg = lambda eps, inner: (lambda P: rg.agents.EpsilonGreedy(
        rg.Configuration({**rg.agents.epsilon_greedy_args, "num_products": P, "epsilon": eps, "random_seed": 3}), inner(P)))
    run("organic_count ALL_DATA dense", 200, 8, oc, T.ALL_DATA, 300, 2000, 3)
    run("bandit_count ALL_DATA dense", 200, 8, bc, T.ALL_DATA, 300, 2000, 3)
    run("bandit_count LAST_STEP dense", 200, 8, bc, T.LAST_STEP, 300, 2000, 3)
    run("bandit_count SLIDING_WINDOW_ALL_DATA dense", 100, 6, bc, T.SLIDING_WINDOW_ALL_DATA, 100, 1500, 3, window=7)
    run("eps(0.2) bandit_count ALL_EXPLORATION_DATA dense", 100, 6, eg(0.2, bc), T.ALL_EXPLORATION_DATA, 100, 1500, 3)
    run("eps(0.02) bandit_count ALL_EXPLORATION_DATA dense (shards without a kept call)", 100, 6, eg(0.02, bc),
        T.ALL_EXPLORATION_DATA, 50, 40, 4)
    run("eps(0.2) organic_count SLIDING_WINDOW_EXPLORATION_DATA dense", 100, 6, eg(0.2, oc), T.SLIDING_WINDOW_EXPLORATION_DATA,
        100, 1500, 3, window=5)
    # (the sparse acting table resolves a row from its first 64 cells: catalogues where a few thousand users do not
    # saturate whole rows, as at P = 100 000)
    run("bandit_count ALL_DATA sparse", 3000, 8, bc, T.ALL_DATA, 300, 1500, 3, sparse_min=1024)
    run("organic_count ALL_DATA sparse", 3000, 8, oc, T.ALL_DATA, 300, 1500, 3, sparse_min=1024)
    run("eps(0.2) bandit_count ALL_EXPLORATION_DATA sparse", 3000, 8, eg(0.2, bc), T.ALL_EXPLORATION_DATA, 300, 1500, 3, sparse_min=1024)
    run("organic_count ALL_DATA dense C3 shape", 10000, 20, oc, T.ALL_DATA, 1000, 100000, 2)
    ok = all(c["equal"] for c in cases)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "sharded evolution == single-GPU evolution", "world": world,
                          "all_equal_on_every_rank": bool(flag.item()), "cases": cases}))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- simulated user-steps/sec of the reco-gym-v1 hot path on N B200s (driver contract).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU baseline on the box's host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, weak scaling

Default workload `c3` (BASELINE.json configs[2], the one the metric "reco-env-v1, P=10k" is quoted on):
num_products=10000, K=20, 1M users per GPU per step, organic_count agent (trained with the
reference-compatible OrganicCount.train on logged users, compiled to a device table).
A step = one fused rollout of `users_per_gpu` fresh users to the end of their episodes on every
rank + the all-reduce of the CTR counters.  1 user-step = one Markov event of one user
(SURVEY.md 8d).  Prints ONE JSON line on rank 0: `value` = device-resident throughput (CUDA events on the launching
stream, L2 flushed between steps), `e2e` = the same through the host C ABI (`rg_rollout_host`: table H2D, counters D2H
inside the timed region), `roofline` (HBM, as BASELINE.json asks; the fused rollout keeps user state on chip),
`roofline_mufu` / `roofline_fp32` (the units that actually bound the kernel), `cpu_baseline` (the C oracle port of the
reference algorithm on all host cores, bounded sample), `clocks`, `gpu_launches`.
Other workloads (`--workload`): c2 (P=100, K=5, uniform logging policy), c4 (P=100 000, K=64, bandit_count table),
c5 (A/B sweep: RandomAgent, OrganicCount, BanditCount, OrganicUserEventCounter, BanditMFSquare over the same users).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (P, K, policy, users per GPU per step)
    "c2": dict(P=100, K=5, policy="uniform", users=1_000_000,
               label="reco-env-v1 P=100 K=5 1M users/GPU/step, default uniform logging policy (configs[1])"),
    "c3": dict(P=10000, K=20, policy="organic_count", users=1_000_000,
               label="reco-env-v1 P=10000 K=20 1M users/GPU/step, organic_count agent (configs[2])"),
    "c4": dict(P=100000, K=64, policy="bandit_count", users=1_250_000,
               label="reco-env-v1 P=100000 K=64 1.25M users/GPU/step (10M over 8 GPUs), bandit_count agent (configs[3])"),
    # verify_agents-style A/B sweep (evaluate_agent.py:732-750): every agent is rolled out over the SAME users
    "c5": dict(P=10000, K=20, policy=["random", "organic_count", "bandit_count", "organic_user_count", "bandit_mf"],
               users=200_000,
               label="A/B sweep: 5 leaderboard agents x 200k users/GPU/step, P=10000 K=20 (configs[4]: 1M users per "
                     "agent = 5 steps at 1 GPU, 1 step at 5+ GPUs)"),
}
SEED = 42
# algorithmic HBM bytes per user-step of the step-synchronous protocol
# (4K read omega + p_o*4K write + (1-f_org)*(4K+4) beta gather + 12 state/action/event bytes)
ALG_BYTES_PER_STEP = {5: 55.0, 20: 175.0, 64: 520.0}
F_ORG = 0.23  # fraction of events that are organic views


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def measured_traffic(user_steps_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the rollout kernel from the committed ncu --set full
    capture (profiles/r01_roofline_traffic.json), scaled from the captured launch to this launch's size."""
    try:
        prof = os.path.join(ROOT, "profiles")
        name = "r02_roofline_traffic.json" if os.path.exists(os.path.join(prof, "r02_roofline_traffic.json")) else "r01_roofline_traffic.json"
        t = json.load(open(os.path.join(prof, name)))
        per_step = (t["dram_bytes_read"] + t["dram_bytes_write"]) / float(t["user_steps_in_captured_launch"])
        return per_step * user_steps_per_launch
    except Exception:
        return None


def static_params(P, K):
    import recogym_b200 as rg
    from recogym_b200.envs.static_params import draw_static_params
    cfg = rg.Configuration({**rg.env_1_args, "num_products": P, "K": K, "random_seed": SEED})
    return cfg, draw_static_params(cfg)


def build_policies(kinds, P, K, n_train, device=0):
    """-> [(kind, device policy name, table or None)].  Trainable agents are trained exactly as test_agent does
    (bench_agents.py:169-190) on the SAME n_train users logged under the default uniform policy: the count agents by
    the device training path (fused rollout -> ordered log -> scatter-add -> arg-max table; recogym_b200/logs.py),
    BanditMFSquare through the RecoEnv1 facade (its train() is torch code), then compiled to device policies
    (recogym_b200/policy.py).  Everything runs on CUDA device `device` (the rank's own GPU)."""
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.logs import generate_logs_device
    A = rg.agents
    agents = {}
    for kind in kinds:
        if kind == "organic_count":
            agents[kind] = A.OrganicCount(rg.Configuration({**A.organic_count_args, "num_products": P}))
        elif kind == "bandit_count":
            agents[kind] = A.BanditCount(rg.Configuration({**A.bandit_count_args, "num_products": P}))
        elif kind == "bandit_mf":
            import torch
            torch.manual_seed(SEED)
            agents[kind] = A.BanditMFSquare(rg.Configuration({**A.bandit_mf_square_args, "num_products": P}))
        elif kind == "organic_user_count":
            agents[kind] = A.OrganicUserEventCounterAgent(rg.Configuration(
                {**A.organic_user_count_args, "num_products": P, "random_seed": SEED}))
        elif kind in ("uniform", "random"):
            agents[kind] = A.RandomAgent(rg.Configuration({**A.random_args, "num_products": P, "random_seed": SEED}))
        else:
            raise ValueError(kind)
    count_agents = [a for k, a in agents.items() if k in ("organic_count", "bandit_count")]
    facade_agents = [a for k, a in agents.items() if k == "bandit_mf"]
    if count_agents or facade_agents:
        env = rg.make("reco-gym-v1")
        env.init_gym({**rg.env_1_args, "num_products": P, "K": K, "random_seed": SEED, "device": device})
        for a in count_agents:
            prev = a.device_last_view() if hasattr(a, "device_last_view") else -1
            log = generate_logs_device(env.device_env, n_train, "uniform", None, None, 0, 0, L.MATH_TENSOR,
                                       prev_last_view=prev, train_mode=L.TRAIN_TEST_AGENT)
            a.device_train(log)
            del log
        n_facade = min(n_train, 200)      # one launch per gym step: the facade is the compatibility path, not the fast one
        for uid in range(n_facade if facade_agents else 0):
            env.reset(uid)
            new_obs, _, done, reward = env.step(None)
            while not done:
                old_obs = new_obs
                action, new_obs, reward, done, _ = env.step_offline(old_obs, reward, done)
                for a in facade_agents:
                    a.train(old_obs, action, reward, False)
            old_obs = new_obs
            action, _, reward, done, _ = env.step_offline(old_obs, reward, done)
            for a in facade_agents:
                a.train(old_obs, action, reward, True)
    out = []
    for kind in kinds:
        pol = agents[kind].device_policy()
        table = pol.get("table")
        if table is not None and hasattr(table, "cpu"):
            table = table.cpu().numpy()
        out.append((kind, pol["policy"], table))
    return out


def host_info():
    """CPU model / cores / load of the box the CPU arm runs on (the same 128-thread sample swung 6x between two hosts
    in round 1: the line now says what it ran on)."""
    model = "?"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    try:
        nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except Exception:
        nodes = None
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = os.cpu_count()
    try:
        load = [round(x, 1) for x in os.getloadavg()]
    except Exception:
        load = None
    return {"cpu_model": model, "logical_cpus": os.cpu_count(), "affinity": aff, "numa_nodes": nodes, "loadavg": load}


def cpu_policy(wl, n_train):
    """The acting policy of the CPU arm == the GPU arm's: the count agents' tables are recomputed from the ORACLE's
    own log of the same n_train training users (oracle/train_oracle.py), so that no kernel of ours is on the CPU arm's
    path; bench.py checks them against the device-trained tables when both exist."""
    from oracle import oracle as O
    from oracle import train_oracle as T
    kinds = wl["policy"] if isinstance(wl["policy"], list) else [wl["policy"]]
    kind = "organic_user_count" if "organic_user_count" in kinds else kinds[0]
    if kind in ("uniform", "random"):
        return "uniform", None
    if kind == "organic_user_count":
        return "user_views", None
    cfg, sp = static_params(wl["P"], wl["K"])
    oe = O.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                     cfg.sigma_omega_initial, cfg.sigma_omega, SEED)
    log = O.rollout(oe, 0, n_train, policy="uniform", log=True, threads=os.cpu_count() or 1)["log"]
    return "table", T.sparse_count_table(kind, wl["P"], log, 0)


def cpu_reference_sample(wl, policy, table, threads, seconds, user0=0):
    """Times the CPU oracle port of the reference algorithm on `threads` host threads for about `seconds` on users
    user0.. of the workload; -> (user-steps/s, users simulated, seconds)."""
    from oracle import oracle as O
    cfg, sp = static_params(wl["P"], wl["K"])
    oe = O.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                     cfg.sigma_omega_initial, cfg.sigma_omega, SEED)
    chunk = max(1, {100: 4096, 10000: 64, 100000: 8}.get(wl["P"], 64) * threads // (1 if threads > 1 else 16))
    n_cpu, ev_cpu, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        rc = O.rollout(oe, user0 + n_cpu, chunk, policy=policy, table=table, threads=threads)
        n_cpu += chunk
        ev_cpu += int(rc["totals"][1])
    dt = time.perf_counter() - t0
    return ev_cpu / dt, n_cpu, dt


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, wl):
    """--impl reference: the reference's algorithm on the box's host cores.  The reference itself is
    Python and does not travel to the GPU box, so this is the C oracle port (oracle/rg_oracle.c),
    all host threads, each step a bounded sample (a few seconds) of the same workload: same parameters, same users
    0.., same acting policy (trained table) as the GPU arm and its `cpu_baseline` leg."""
    cores = os.cpu_count() or 1
    policy, table = cpu_policy(wl, args.train_users)
    secs = 4.0
    for w in range(args.warmup):
        cpu_reference_sample(wl, policy, table, cores, 1.0, user0=50_000_000)
    events_s, users, dt_all = [], 0, 0.0
    for s in range(args.steps):
        v, n, dt = cpu_reference_sample(wl, policy, table, cores, secs, user0=users)
        events_s.append(v * dt); users += n; dt_all += dt
    val = sum(events_s) / dt_all
    one, n1, dt1 = cpu_reference_sample(wl, policy, table, 1, 4.0)
    sample = "users 0..%d of the same workload, %d steps of ~%.0f s, trained table of the GPU arm" % (users - 1, args.steps, secs)
    line = {
        "impl": "reference", "metric": "simulated user-steps/sec", "value": val, "unit": "user-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "users_timed": users, "seed": SEED, "host": host_info()},
        "cpu_baseline": {"value": val, "unit": "user-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "one_core": {"value": one, "users": n1}},
        "e2e": {"value": val, "unit": "user-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--users", type=int, default=0, help="users per GPU per step (default: workload's)")
    ap.add_argument("--math", default="tensor", choices=["fast", "exact", "tensor"])
    ap.add_argument("--train-users", type=int, default=1000, help="logged users the agents are trained on (SURVEY 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.users:
        wl["users"] = args.users
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.parallel import allreduce_counters

    if L.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device and recogym_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    math_mode = {"fast": L.MATH_FAST, "exact": L.MATH_EXACT, "tensor": L.MATH_TENSOR}[args.math]
    P, K, U = wl["P"], wl["K"], wl["users"]

    cfg, sp = static_params(P, K)
    kinds = wl["policy"] if isinstance(wl["policy"], list) else [wl["policy"]]
    policies = build_policies(kinds, P, K, args.train_users, device=local_rank)
    de = rg.DeviceEnv(sp, cfg.sigma_omega_initial, cfg.sigma_omega, SEED, device=local_rank)
    tables_dev = [torch.as_tensor(t, device=dev) if t is not None else None for _, _, t in policies]

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    job_tot = torch.zeros(16, dtype=torch.int64, device=dev)

    def step(s, tot):
        """one fused rollout of U fresh users on this rank (every agent of an A/B sweep sees the same users); the
        counters accumulate on the device"""
        first = (s * world + rank) * U
        for (_, pol, _), tdev in zip(policies, tables_dev):
            de.rollout_device(U, pol, tdev, first, 0, math_mode, per_user=False, totals=tot)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm_tot = torch.zeros(16, dtype=torch.int64, device=dev)
    for w in range(args.warmup):
        step(1000 + w, warm_tot)
    allreduce_counters(warm_tot)          # warms the collective up as well
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The timed region: K steps (L2 flushed before each) + ONE all-reduce of the job's CTR counters at the end -- the
    # only collective of the path.  Ranks are not synchronised between steps (round 1 all-reduced after
    # every step, which made every rank wait for the slowest launch tail 8 times per second).
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)              # L2 flush between timed iterations (40 us of 256 MiB writes, inside the bracket)
        evs[s][0].record()
        step(s, job_tot)
        evs[s][1].record()
    allreduce_counters(job_tot)
    ev1.record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_dev = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    clocks = sampler.stop() if rank == 0 else None
    ev = job_tot.cpu().numpy()            # all-reduced: job-wide counters of the timed steps
    events = int(ev[1])
    value = events / (total_ms * 1e-3)

    # ---- rank-count invariance: the counters of a fixed user range do not depend on how it is sharded
    rank_invariant = None
    if world > 1:
        from recogym_b200.parallel import shard_range
        n_fix = 1_000_000 if P <= 10000 else 100_000
        first, count = shard_range(n_fix, rank, world, 7_000_000)
        part = de.rollout_device(count, policies[0][1], tables_dev[0], first, 0, math_mode, per_user=False, hist=True)
        red = [part["totals"], part["view_hist"], part["action_hist"], part["click_hist"]]
        allreduce_counters(*red)
        if rank == 0:
            whole = de.rollout_device(n_fix, policies[0][1], tables_dev[0], 7_000_000, 0, math_mode, per_user=False, hist=True)
            rank_invariant = bool(torch.equal(whole["totals"][:10], part["totals"][:10]) and
                                  all(torch.equal(whole[k], part[k]) for k in ("view_hist", "action_hist", "click_hist")))
        barrier()

    # ---- e2e: the reference-facing host API (C ABI with HOST buffers), copies inside the timed region
    for _, pol, tab in policies:          # untimed: grows the env's staging arena once
        de.rollout_host(1024, pol, tab, 9_000_000, 0, math_mode, per_user=False)
    barrier()
    t0 = time.perf_counter()
    e2e_events = 0
    d2h = 0
    for s in range(args.steps):
        first = ((s + 5000) * world + rank) * U
        for _, pol, tab in policies:
            r = de.rollout_host(U, pol, tab, first, 0, math_mode, per_user=False)
            e2e_events += int(r["totals"][1])
        d2h = 128 * len(policies)
    if world > 1:
        # counters of the host path are reduced with the same collective, on device
        tt = torch.tensor([e2e_events], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        e2e_events = int(tt.item())
    barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = e2e_events / float(e2e_dt.item())
    h2d = sum(4 * P for _, _, t in policies if t is not None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src, sm_max = peaks()
    per_gpu_steps_s = value / world
    alg_bytes = ALG_BYTES_PER_STEP.get(K, 4.0 * K * 2.2 + 16)
    avg_launch_s = (total_ms * 1e-3) / args.steps / len(policies)
    launch_events = events / world / args.steps / len(policies)
    achieved_gbs = alg_bytes * launch_events / avg_launch_s / 1e9
    # compute roofline (the physically binding one): fp32 FMA pipe of the organic contraction
    flops_per_step = F_ORG * (2.0 * P * K)
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6
    line = {
        "metric": "simulated user-steps/sec", "value": value, "unit": "user-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"tensor": "f64+f16x2", "fast": "f64+f32", "exact": "f64"}[args.math],
        "data": "synthetic",
        "config": {"workload": wl["label"], "num_products": P, "K": K, "users_per_gpu_per_step": U,
                   "policy": wl["policy"], "device_policies": [p for _, p, _ in policies], "math_mode": args.math, "seed": SEED, "parallelism": "users sharded dp%d" % world,
                   "l2": "flushed between timed iterations (256 MiB write); Gamma is L2-resident by design within a launch",
                   "train_users": args.train_users, "step_ms_min_max": [round(min(step_ms), 3), round(max(step_ms), 3)],
                   "rank_invariant_counters": rank_invariant,
                   "events_timed": events, "ctr": float(ev[4]) / max(1.0, float(ev[3])),
                   "fast_path_fallbacks": int(ev[6]), "tier3_resamples": int(ev[8]), "f64_ambiguous": int(ev[7]),
                   "pass1_cycles_per_128_draws": round(float(ev[11]) / max(1.0, float(ev[2])) * 128.0, 1),
                   "phase_cycle_share": dict(zip(("compact", "pass1", "pass2", "resample", "advance", "omega"),
                                                 [round(float(x) / max(1.0, float(ev[10:16].sum())), 4) for x in ev[10:16]]))},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "user-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": args.steps * len(policies) * (5 if (args.math != "exact" and U >= 4 * 296 * 128) else 1),  # rollout (+ 4 schedule kernels)
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": measured_traffic(launch_events), "peak_source": peak_src,
                     "algorithmic_bytes_per_user_step": alg_bytes,
                     "note": "fused rollout keeps user state on chip: HBM is not the physical bound (SURVEY.md 8d); "
                             "see roofline_fp32 / roofline_mufu"},
    }
    organic_per_s = float(ev[2]) / (total_ms * 1e-3) / world          # organic draws per second per GPU
    if args.math == "tensor":
        # tensor mode: the contraction is on tcgen05; the binding unit of pass 1 is the MUFU pipe
        # (one ex2 per product per organic draw; 16 ex2/clk/SM)
        mufu_peak = 148 * 15.8 * sm_max * 1e6     # measured: scripts/micro/mufu_rate.cu, profiles/r01_mufu_rate.txt (nominal 16)
        line["roofline_mufu"] = {"bound": "mufu-ex2", "achieved": organic_per_s * P / 1e12, "peak": mufu_peak / 1e12,
                                 "unit": "Tex2/s", "frac": organic_per_s * P / mufu_peak,
                                 "ex2_per_organic_draw": P}
        line["roofline_tensor"] = {"bound": "tensor", "achieved": organic_per_s * 3 * 2.0 * P * (K + 2) / 1e12,
                                   "peak": None, "unit": "TFLOP/s (fp16 hi/lo operands, three-term split, fp32 accumulate)",
                                   "note": "not the binding unit: 22% tensor-pipe active at the benched launch (profiles/r02_tensor_1M_ncu_summary.txt)"}
    else:
        line["roofline_fp32"] = {"bound": "fp32-fma", "achieved": per_gpu_steps_s * flops_per_step / 1e12,
                                 "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                                 "frac": per_gpu_steps_s * flops_per_step / fp32_peak,
                                 "flops_per_user_step": flops_per_step}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpol, ctab = cpu_policy(wl, args.train_users)
        if ctab is not None and policies[0][2] is not None and cpol == "table" and policies[0][1] == "table":
            # the table the oracle derives from ITS log of the training users == the device-trained table
            line["config"]["device_trained_table_equals_oracle"] = bool(np.array_equal(ctab, policies[0][2]))
        v_all, n_all, dt_all = cpu_reference_sample(wl, cpol, ctab, cores, 12.0)
        v_one, n_one, dt_one = cpu_reference_sample(wl, cpol, ctab, 1, 5.0)
        line["cpu_baseline"] = {"value": v_all, "unit": "user-steps/s", "cores": cores, "kind": "port",
                                "sample": "users 0..%d of the same workload, same trained table (%.1f s; C oracle port of "
                                          "the reference algorithm, oracle/rg_oracle.c, %d host threads)" % (n_all - 1, dt_all, cores),
                                "one_core": {"value": v_one, "users": n_one, "seconds": round(dt_one, 1)},
                                "host": host_info()}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

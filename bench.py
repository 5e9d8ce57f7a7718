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
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")))
        per_step = (t["dram_bytes_read"] + t["dram_bytes_write"]) / float(t["user_steps_in_captured_launch"])
        return per_step * user_steps_per_launch
    except Exception:
        return None


def static_params(P, K):
    import recogym_b200 as rg
    from recogym_b200.envs.static_params import draw_static_params
    cfg = rg.Configuration({**rg.env_1_args, "num_products": P, "K": K, "random_seed": SEED})
    return cfg, draw_static_params(cfg)


def build_policies(kinds, P, K, n_train):
    """-> [(kind, device policy name, table or None)].  Trainable agents are trained exactly as test_agent does
    (bench_agents.py:176-190) on the SAME n_train logged users through the RecoEnv1 facade (GPU, default uniform
    logging policy), then compiled to device policies (recogym_b200/policy.py)."""
    import recogym_b200 as rg
    A = rg.agents
    agents = {}
    for kind in kinds:
        if kind == "organic_count":
            agents[kind] = A.OrganicCount(rg.Configuration({**A.organic_count_args, "num_products": P}))
        elif kind == "bandit_count":
            agents[kind] = _SparseBanditCount(P)
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
    trainable = [a for k, a in agents.items() if k in ("organic_count", "bandit_count", "bandit_mf")]
    if trainable:
        env = rg.make("reco-gym-v1")
        env.init_gym({**rg.env_1_args, "num_products": P, "K": K, "random_seed": SEED})
        for uid in range(n_train):
            env.reset(uid)
            new_obs, _, done, reward = env.step(None)
            while not done:
                old_obs = new_obs
                action, new_obs, reward, done, _ = env.step_offline(old_obs, reward, done)
                for a in trainable:
                    a.train(old_obs, action, reward, False)
            old_obs = new_obs
            action, _, reward, done, _ = env.step_offline(old_obs, reward, done)
            for a in trainable:
                a.train(old_obs, action, reward, True)
    out = []
    for kind in kinds:
        pol = agents[kind].device_policy()
        out.append((kind, pol["policy"], pol.get("table")))
    return out


class _SparseBanditCount:
    """BanditCount (agents/bandit_count.py) with the P x P matrices held sparsely: dense is
    3 x 80 GB at P=100k.  Same update rule, same argmax with the (0+1)/(0+2) prior."""

    def __init__(self, P):
        self.P, self.pulls, self.clicks, self.lpv = P, {}, {}, None

    def train(self, observation, action, reward, done=False):
        if action is not None and reward is not None:
            ix, jx = self.lpv, int(action["a"])
            if observation.sessions():
                self.lpv = observation.sessions()[-1]["v"]
            if ix is not None:
                self.pulls[(ix, jx)] = self.pulls.get((ix, jx), 0) + 1
                self.clicks[(ix, jx)] = self.clicks.get((ix, jx), 0) + reward

    def device_policy(self):
        table = np.zeros(self.P, np.int32)
        best = {}
        for (ix, jx), n in self.pulls.items():
            ctr = (self.clicks[(ix, jx)] + 1) / (n + 2)
            cur = best.get(ix)
            if ctr > 0.5 and (cur is None or ctr > cur[0] or (ctr == cur[0] and jx < cur[1])):
                best[ix] = (ctr, jx)
        for ix, (_, jx) in best.items():
            table[ix] = jx
        return {"policy": "table", "table": table}


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
    all host threads, each step a bounded sample of the same workload."""
    from oracle import oracle as O
    cfg, sp = static_params(wl["P"], wl["K"])
    oe = O.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                     cfg.sigma_omega_initial, cfg.sigma_omega, SEED)
    cores = os.cpu_count() or 1
    policy, table = ("uniform", None)
    if isinstance(wl["policy"], list):
        policy = "user_views"      # the most expensive acting policy of the sweep on the CPU (dense P-vector per act)
    elif wl["policy"] != "uniform":
        policy, table = "table", np.random.RandomState(SEED).randint(0, wl["P"], wl["P"]).astype(np.int32)
    users = {100: 8192, 10000: 96, 100000: 12}.get(wl["P"], 96) * cores   # a few seconds per step on all cores
    base = 10_000_000
    for w in range(args.warmup):
        O.rollout(oe, base + w * users, max(cores, users // 8), policy=policy, table=table, threads=cores)
    events = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        r = O.rollout(oe, s * users, users, policy=policy, table=table, threads=cores)
        events += int(r["totals"][1])
    dt = time.perf_counter() - t0
    val = events / dt
    sample = "%d users/step x %d steps of the same workload, users 0.." % (users, args.steps)
    line = {
        "impl": "reference", "metric": "simulated user-steps/sec", "value": val, "unit": "user-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "users_per_step": users, "seed": SEED},
        "cpu_baseline": {"value": val, "unit": "user-steps/s", "cores": cores, "kind": "port", "sample": sample},
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
    ap.add_argument("--train-users", type=int, default=200)
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
    policies = build_policies(kinds, P, K, args.train_users)
    de = rg.DeviceEnv(sp, cfg.sigma_omega_initial, cfg.sigma_omega, SEED, device=local_rank)
    tables_dev = [torch.as_tensor(t, device=dev) if t is not None else None for _, _, t in policies]
    policy, table = policies[0][1], policies[0][2]          # cpu_baseline leg: the first policy of the workload

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    step_tots = torch.zeros((args.steps, 16), dtype=torch.int64, device=dev)

    def step(s, tot):
        """one fused rollout of U fresh users on this rank + the all-reduce of the CTR counters"""
        first = (s * world + rank) * U
        for (_, pol, _), tdev in zip(policies, tables_dev):   # every agent of an A/B sweep sees the same users
            de.rollout_device(U, pol, tdev, first, 0, math_mode, per_user=False, totals=tot)
        allreduce_counters(tot)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(1000 + w, torch.zeros(16, dtype=torch.int64, device=dev))
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)              # L2 flush between timed iterations (outside the event pair)
        evs[s][0].record()
        step(s, step_tots[s])
        evs[s][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    clocks = sampler.stop() if rank == 0 else None
    ev = step_tots.sum(0).cpu().numpy()   # every row was all-reduced: job-wide counters of the timed steps
    events = int(ev[1])
    value = events / (total_ms * 1e-3)

    # ---- e2e: the reference-facing host API (C ABI with HOST buffers), copies inside the timed region
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"tensor": "f64+tf32x2", "fast": "f64+f32", "exact": "f64"}[args.math],
        "data": "synthetic",
        "config": {"workload": wl["label"], "num_products": P, "K": K, "users_per_gpu_per_step": U,
                   "policy": wl["policy"], "device_policies": [p for _, p, _ in policies], "math_mode": args.math, "seed": SEED, "parallelism": "users sharded dp%d" % world,
                   "l2": "flushed between timed iterations (256 MiB write); Gamma is L2-resident by design within a launch",
                   "events_timed": events, "ctr": float(ev[4]) / max(1.0, float(ev[3])),
                   "fast_path_fallbacks": int(ev[6]), "tier3_resamples": int(ev[8]), "f64_ambiguous": int(ev[7]),
                   "pass1_cycles_per_128_draws": round(float(ev[11]) / max(1.0, float(ev[2])) * 128.0, 1),
                   "phase_cycle_share": dict(zip(("compact", "pass1", "pass2", "resample", "advance", "omega"),
                                                 [round(float(x) / max(1.0, float(ev[10:16].sum())), 4) for x in ev[10:16]]))},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "user-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": args.steps * len(policies) * (4 if (args.math != "exact" and U >= 4 * 296 * 128) else 1),  # rollout (+ 3 schedule kernels)
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
                                   "peak": None, "unit": "TFLOP/s (tf32, three-term split)",
                                   "note": "not the binding unit: <20% tensor-pipe active (profiles/)"}
    else:
        line["roofline_fp32"] = {"bound": "fp32-fma", "achieved": per_gpu_steps_s * flops_per_step / 1e12,
                                 "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                                 "frac": per_gpu_steps_s * flops_per_step / fp32_peak,
                                 "flops_per_user_step": flops_per_step}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        oe = O.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                         cfg.sigma_omega_initial, cfg.sigma_omega, SEED)
        cores = os.cpu_count() or 1
        chunk = {100: 4096, 10000: 64, 100000: 8}.get(P, 64) * cores   # bounded sample: stop after ~12 s
        n_cpu, ev_cpu, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0:
            rc = O.rollout(oe, n_cpu, chunk, policy=policy, table=table, threads=cores)
            n_cpu += chunk
            ev_cpu += int(rc["totals"][1])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ev_cpu / dt, "unit": "user-steps/s", "cores": cores, "kind": "port",
                                "sample": "users 0..%d of the same workload (%.1f s; C oracle port of the reference "
                                "algorithm, oracle/rg_oracle.c, %d host threads)" % (n_cpu - 1, dt, cores)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Headline benchmark: env-steps/sec of a batched random-action rollout (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 2, per GPU): 2^18 copies of benchmark puzzle `level1/Road Block`
(8x5 grid incl. border, 3 objects), T = 1024 uniformly random actions per environment
(torch.randint on the device, seed 1234 + rank), no reset on termination (reference semantics).
One bench "step" = pw_reset + ONE pw_rollout launch = 2^18 * 1024 env-steps per GPU, writing
float32 reward[T][N] and uint8 terminated[T][N].  Weak scaling: every GPU gets its own 2^18
environments; the only collective is one NCCL all-reduce of the int64[4] counter vector per
rollout.  Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md section 6).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "env-steps/sec (batched random-action rollout)"
UNIT = "env-steps/s"
PUZZLE_KEY = "level1/Road Block"
N_ENVS = 1 << 18
T_STEPS = 1024
ACTION_SEED = 1234


def _workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"cfg2: 2^18 copies of {PUZZLE_KEY} per GPU, T=1024 random actions, "
        "reward f32 + terminated u8 written per step",
        "puzzle": PUZZLE_KEY,
        "envs_per_gpu": N_ENVS,
        "rollout_steps": T_STEPS,
        "parallelism": f"shard{n_gpus}",
        "l2": "inputs (256 MiB actions) and outputs (1.25 GiB) per step exceed the 126 MB L2",
    }


# ---------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []  # (arrival time, csv line)
        self.proc = None
        self.thread = None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows
                  if self.t_begin is not None and self.t_begin <= t <= (self.t_end or t) + 0.05]
        window = "timed region"
        if len(inside) < 3:  # very short timed region: use everything since the warm-up started
            inside, window = [r for _, r in self.rows], "warm-up + timed region"
        for row in inside:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "window": window,
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------- CPU baseline
def cpu_port_throughput(target_seconds: float, threads: int = 0):
    """Times oracle/pw_oracle.c (the reference's table-driven algorithm restated in C, OpenMP
    over environments) on a bounded sample of the SAME workload.  Returns (env-steps/s,
    threads, sample description).  This is the one place bench.py executes oracle/."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import numpy as np

    import c_oracle
    from pushworld_b200 import data

    puzzle = c_oracle.COraclePuzzle(data.puzzle_text(PUZZLE_KEY))
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it)
    threads = threads or len(os.sched_getaffinity(0))
    rng = np.random.default_rng(ACTION_SEED)

    def run(n_envs, n_steps):
        init = np.array(puzzle.py.initial_state, dtype=np.int32)
        x = np.ascontiguousarray(np.repeat(init[:, 0:1], n_envs, axis=1))
        y = np.ascontiguousarray(np.repeat(init[:, 1:2], n_envs, axis=1))
        actions = rng.integers(0, 4, size=(n_steps, n_envs), dtype=np.uint8)
        reward = np.empty((n_steps, n_envs), np.float32)
        term = np.empty((n_steps, n_envs), np.uint8)
        t0 = time.perf_counter()
        done = puzzle.rollout(x, y, actions, reward, term, threads=threads)
        return done / (time.perf_counter() - t0), done

    rate, _ = run(1 << 12, 256)  # calibration (also warms the thread pool)
    n_steps = 256
    n_envs = max(1 << 12, int(rate * target_seconds / n_steps))
    n_envs = min(n_envs, 1 << 22)
    rate, done = run(n_envs, n_steps)
    return rate, threads, f"{n_envs} envs x {n_steps} steps of {PUZZLE_KEY} ({done} env-steps)"


def python_port_throughput(seconds: float = 3.0):
    """Single-core rate of the pure-Python restatement (closest stand-in for the reference's own
    Python get_next_state, which cannot travel to the GPU box)."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import random

    import pushworld_oracle as po
    from pushworld_b200 import data

    puzzle = po.OraclePuzzle(data.puzzle_text(PUZZLE_KEY))
    rng = random.Random(ACTION_SEED)
    state, n = puzzle.initial_state, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(500):
            nxt = puzzle.get_next_state(state, rng.randrange(4))
            po.step_reward(puzzle, state, nxt)
            state = nxt
        n += 500
    return n / (time.perf_counter() - t0)


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # W warm-up + K timed steps; each step is a bounded sample of ~2 s of CPU work
    for _ in range(max(args.warmup, 1)):
        cpu_port_throughput(0.3)
    rates, t0 = [], time.perf_counter()
    sample = ""
    threads = 0
    for _ in range(args.steps):
        rate, threads, sample = cpu_port_throughput(2.0)
        rates.append(rate)
    elapsed = time.perf_counter() - t0
    value = statistics.mean(rates)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32",
        "data": "synthetic",
        "config": _workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + " per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference arm = oracle/pw_oracle.c (C port of the reference's table-driven "
        "get_next_state; the Python reference cannot travel to the GPU box and its C++ twin "
        "needs Boost), OpenMP over all host threads",
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ---------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import torch

    from pushworld_b200 import data, distributed
    from pushworld_b200.vector_env import PushWorldVectorEnv, RolloutOutput

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    rank, local_rank, world = distributed.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    env = PushWorldVectorEnv([data.puzzle_text(PUZZLE_KEY)], N_ENVS, puzzle_ids=0, device=dev)
    n_obj = env.engine.n_max
    gen = torch.Generator(device=dev).manual_seed(ACTION_SEED + rank)
    actions = torch.randint(0, 4, (T_STEPS, N_ENVS), dtype=torch.uint8, device=dev, generator=gen)
    out = RolloutOutput(
        reward=torch.empty((T_STEPS, N_ENVS), dtype=torch.float32, device=dev),
        terminated=torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8, device=dev),
        counters=torch.zeros(4, dtype=torch.int64, device=dev),
    )

    def one_step(ev0=None, ev1=None):
        env.reset()
        if ev0 is not None:
            ev0.record()
        env.rollout(actions, out=out)
        if ev1 is not None:
            ev1.record()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs ------------------------------------------
    out.counters.zero_()
    kernel_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    distributed.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    start.record()
    for k in range(args.steps):
        one_step(*kernel_events[k])
    totals = out.counters.clone()
    distributed.allreduce_counters(totals)  # the path's single collective (NCCL, int64[4])
    stop.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    distributed.barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = distributed.max_over_ranks(start.elapsed_time(stop), dev)
    total_env_steps = int(totals[0].item())
    assert total_env_steps == world * args.steps * N_ENVS * T_STEPS, total_env_steps
    value = total_env_steps / (elapsed_ms * 1e-3)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kernel_events)

    # ---- end to end: pinned host buffers in, pinned host buffers out ---------------------
    h_act = torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8).pin_memory()
    h_act.copy_(actions)
    h_rew = torch.empty((T_STEPS, N_ENVS), dtype=torch.float32).pin_memory()
    h_term = torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8).pin_memory()
    e2e_steps = min(args.steps, 40)  # ~25 ms each (PCIe bound)

    def e2e_step():
        env.reset()
        env.rollout_host(h_act, h_rew, h_term, chunk_steps=64)

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    distributed.barrier()
    torch.cuda.synchronize()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(e2e_steps):
        e2e_step()
    t_stop.record()
    torch.cuda.synchronize()
    e2e_ms = distributed.max_over_ranks(t_start.elapsed_time(t_stop), dev)
    e2e_value = world * e2e_steps * N_ENVS * T_STEPS / (e2e_ms * 1e-3)
    # the device-resident run and the host-buffer run must produce the same bytes
    assert torch.equal(h_rew[-1], out.reward[-1].cpu()) and torch.equal(h_term[-1], out.terminated[-1].cpu())
    chunks = (T_STEPS + 63) // 64

    # ---- end to end, summaries only: host actions in, per-environment returns out ------
    h_ret = torch.empty(N_ENVS, dtype=torch.float32).pin_memory()
    h_nterm = torch.empty(N_ENVS, dtype=torch.int32).pin_memory()

    def e2e_returns_step():
        env.reset()
        env.rollout_host_returns(h_act, h_ret, h_nterm, chunk_steps=64)

    for _ in range(min(args.warmup, 3)):
        e2e_returns_step()
    distributed.barrier()
    torch.cuda.synchronize()
    r_start, r_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r_start.record()
    for _ in range(e2e_steps):
        e2e_returns_step()
    r_stop.record()
    torch.cuda.synchronize()
    e2e_ret_ms = distributed.max_over_ranks(r_start.elapsed_time(r_stop), dev)
    e2e_ret_value = world * e2e_steps * N_ENVS * T_STEPS / (e2e_ret_ms * 1e-3)
    assert torch.equal(h_nterm.to(torch.int64), out.terminated.sum(dim=0, dtype=torch.int64).cpu())

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (pw_rollout) -----------------------------------
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, TypeError, ValueError):
        pass  # absent or unreadable: the documented fallback
    units = N_ENVS * T_STEPS
    # bytes a T-fused rollout must move per env-step: action 1 + reward 4 + terminated 1, plus the
    # state read once and written once per launch (16 n / T).  SURVEY 8(d)'s 16n+6 is the T=1 figure.
    fused_bpu = 6.0 + 16.0 * n_obj / T_STEPS
    step_bpu = 16.0 * n_obj + 6.0
    achieved = units * fused_bpu / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "k_rollout_ring<3, true> (pw_rollout)",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "peak_source": peak_src,
        "bytes_per_unit": fused_bpu,
        "kernel_ms": kernel_ms,
        "survey_8d_accounting": {
            "bytes_per_unit": step_bpu,
            "achieved": units * step_bpu / (kernel_ms * 1e-3) / 1e9,
            "frac": units * step_bpu / (kernel_ms * 1e-3) / 1e9 / peak,
            "note": "16n+6 B per step is what T separate pw_step launches would move; the fused "
            "rollout keeps the state in registers, so its real traffic is bytes_per_unit above",
        },
    }
    traffic_path = os.path.join(REPO, "profiles", "rollout_traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            traffic = json.load(f)
        roofline["traffic"] = traffic.get("dram_bytes_per_launch")
        # secondary bound: the kernel is bound by the integer ALU pipe, not by HBM
        roofline["secondary_bound"] = traffic.get("secondary_bound")

    cpu_rate, cpu_threads, cpu_sample = cpu_port_throughput(12.0)
    py_rate = python_port_throughput(3.0)

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32",
        "data": "synthetic",
        "config": _workload_config(world),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                         "sample": cpu_sample,
                         "python_port_single_core": py_rate},
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": T_STEPS * N_ENVS,
            "d2h_bytes_per_step": T_STEPS * N_ENVS * 5,
            "ms_per_step": e2e_ms / e2e_steps,
            "api": "PushWorldVectorEnv.rollout_host (pinned host buffers, 3-stream pipeline)",
        },
        "e2e_summaries": {
            "value": e2e_ret_value,
            "unit": UNIT,
            "h2d_bytes_per_step": T_STEPS * N_ENVS,
            "d2h_bytes_per_step": N_ENVS * 8,
            "ms_per_step": e2e_ret_ms / e2e_steps,
            "api": "PushWorldVectorEnv.rollout_host_returns (per-environment return sum + terminated "
            "count come back; per-step arrays stay on the device)",
        },
        "gpu_launches": 2 * args.steps,
        "gpu_launches_e2e": (1 + chunks) * e2e_steps,
        "clocks": clocks,
        "totals": {"env_steps": total_env_steps, "terminated_steps": int(totals[1].item()),
                   "moved_steps": int(totals[2].item()), "goals_gained": int(totals[3].item())},
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


_RESULT_OUT = sys.stdout


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries may write banners to file descriptor 1
    # (NCCL prints its version there when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the
    # whole run and the JSON line goes to a private duplicate of the original stdout.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least three warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        try:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()

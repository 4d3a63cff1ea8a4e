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
ROLLOUT_KERNEL = "k_rollout_pairs4<plain> (pw_rollout on the two-step reachable-state table, four environments per thread)"


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
        self.windows = []  # [begin, end] wall-clock intervals of the timed regions

    def mark_begin(self):
        self.windows.append([time.time(), None])

    def mark_end(self):
        if self.windows and self.windows[-1][1] is None:
            self.windows[-1][1] = time.time()

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
                  if any(b <= t <= (e or t) + 0.05 for b, e in self.windows)]
        window = "timed regions (device-resident + e2e + configs)"
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
def cpu_throughput(target_seconds: float, threads: int = 0, kind: str = "reference"):
    """Times a CPU implementation of the SAME workload on a bounded sample: `kind` "reference" = the
    reference's own C++ getNextState (cpp/src/pushworld_puzzle.cc compiled into oracle/_ref, every
    OpenMP thread stepping its own copy of the puzzle), "port" = oracle/pw_oracle.c (the reference's
    table-driven algorithm restated in C).  Falls back to "port" when oracle/_ref is not available.
    Returns (env-steps/s, threads, sample description, kind).  bench.py executes oracle/ only here."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import numpy as np

    import c_oracle
    import ref_oracle
    from pushworld_b200 import data

    text = data.puzzle_text(PUZZLE_KEY)
    if kind == "reference" and ref_oracle.available():
        puzzle = ref_oracle.RefPuzzle(text)
    else:
        kind = "port"
        puzzle = c_oracle.COraclePuzzle(text)
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
    return rate, threads, f"{n_envs} envs x {n_steps} steps of {PUZZLE_KEY} ({done} env-steps)", kind


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
        cpu_throughput(0.3)
    rates, t0 = [], time.perf_counter()
    sample, kind = "", "reference"
    threads = 0
    for _ in range(args.steps):
        rate, threads, sample, kind = cpu_throughput(2.0)
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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample + " per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference arm = the reference's own C++ getNextState + satisfiesGoal "
        "(cpp/src/pushworld_puzzle.cc, compiled unmodified into oracle/_ref with a 5-function Boost "
        "shim; the Python reference cannot travel to the GPU box), one copy of the puzzle per OpenMP "
        "thread, all host threads" if kind == "reference" else
        "reference arm = oracle/pw_oracle.c (C port of the reference's table-driven get_next_state): "
        "oracle/_ref was not available",
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ---------------------------------------------------------------------------- our arm
def _peak():
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(peaks_path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, TypeError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"  # absent or unreadable: the documented fallback


def _timed(fn, reps, warmup, dev, distributed):
    """Mean ms per call of `fn` over `reps` calls (CUDA events on the current stream, barrier +
    synchronize on both sides, max over ranks)."""
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    distributed.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    distributed.barrier()
    return distributed.max_over_ranks(a.elapsed_time(b), dev) / reps


def run_configs(rank, world, dev, peak, sampler, reps=5):
    """BASELINE configs 3, 4, 5 (+ the reference's real level-0 set): device-timed sub-records.
    Under N GPUs every workload is SHARDED (fixed total size, contiguous slice per rank, strong
    scaling); the rates are whole-job rates (total units / max-over-ranks time)."""
    import torch

    from pushworld_b200 import data, distributed
    from pushworld_b200.engine import Engine
    from pushworld_b200.vector_env import PushWorldVectorEnv, RolloutOutput

    out = {}

    def record(name, workload, kernel, units, ms, fused_bpu, survey_bpu, extra=None):
        rate = units / (ms * 1e-3)
        rec = {"workload": workload, "kernel": kernel, "units": units, "kernel_ms": ms,
               "units_per_s": rate, "n_gpus": world, "sharded": world > 1,
               "bytes_per_unit": fused_bpu, "achieved_gbs": rate * fused_bpu / 1e9,
               "frac": rate * fused_bpu / 1e9 / (peak * world)}
        if survey_bpu is not None:
            rec["survey_8d"] = {"bytes_per_unit": survey_bpu, "achieved_gbs": rate * survey_bpu / 1e9,
                                "frac": rate * survey_bpu / 1e9 / (peak * world)}
        if extra:
            rec.update(extra)
        out[name] = rec

    # ---- config 3: all 223 benchmark puzzles, N = 2^18, T = 1024 ---------------------------
    keys = data.benchmark_keys()
    texts = [data.puzzle_text(k) for k in keys]
    n_total, T = 1 << 18, 1024
    for name, ids in (("cfg3_grouped", None),
                      ("cfg3_interleaved", torch.arange(n_total, dtype=torch.int64) % len(keys))):
        env = PushWorldVectorEnv(texts, n_total, puzzle_ids=ids, device=dev, shard=True)
        n = env.num_envs
        gen = torch.Generator(device=dev).manual_seed(1235 + rank)
        actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device=dev, generator=gen)
        ro = RolloutOutput(reward=torch.empty((T, n), dtype=torch.float32, device=dev),
                           terminated=torch.empty((T, n), dtype=torch.uint8, device=dev),
                           counters=torch.zeros(4, dtype=torch.int64, device=dev))

        def run():
            env.reset()
            env.rollout(actions, out=ro)  # incl. the per-rollout counter all-reduce

        sampler.mark_begin()
        ms = _timed(run, reps, 2, dev, distributed)
        sampler.mark_end()
        n_obj = torch.tensor([p.num_objects for p in env.engine.parsed], dtype=torch.float64)[
            (torch.arange(n_total) % len(keys))].mean().item()
        record(name,
               f"cfg3: 2^18 envs over all {len(keys)} benchmark puzzles "
               f"({'grouped by puzzle' if ids is None else 'env e -> puzzle e mod 223'}), T=1024, "
               "reward f32 + terminated u8 per step",
               "k_rollout_mask (general rule: bounding-box mask prefilter + collision-table look-ups)",
               n_total * T, ms, 6.0 + 16.0 * n_obj / T, 16.0 * n_obj + 6.0, {"unit": "env-steps"})
        del env, actions, ro
    torch.cuda.empty_cache()

    # ---- the reference's real level-0 set (15,400 generated puzzles), RL-style episodes ------
    l0_keys = data.level0_keys()
    eng = Engine([data.level0_text(k) for k in l0_keys], device=dev)
    lo, hi = distributed.shard_range(1 << 18, rank, world)
    n, T0 = hi - lo, 256
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    pid = torch.randint(0, len(l0_keys), (n,), dtype=torch.int32, device=dev, generator=gen)
    actions = torch.randint(0, 4, (T0, n), dtype=torch.uint8, device=dev, generator=gen)
    xy = eng.new_state(n)
    steps = torch.zeros(n, dtype=torch.int32, device=dev)
    episode = torch.zeros(n, dtype=torch.int32, device=dev)
    reward = torch.empty((T0, n), dtype=torch.float32, device=dev)
    term = torch.empty((T0, n), dtype=torch.uint8, device=dev)
    trunc = torch.empty((T0, n), dtype=torch.uint8, device=dev)
    eng.reset(xy, pid, steps=steps)

    def run_level0():
        eng.rollout_episodes(xy, pid, actions, steps, episode, seed=11, max_steps=100,
                             reward=reward, terminated=term, truncated=trunc)

    sampler.mark_begin()
    ms = _timed(run_level0, reps, 2, dev, distributed)
    sampler.mark_end()
    record("level0_real_episodes",
           "all 15,400 puzzles of the reference's benchmark/puzzles/level0.zip in ONE engine, 2^18 envs with "
           "random puzzle ids, pw_rollout_episodes (max_steps=100, puzzle re-sampled on the device after "
           "every episode), T=256, reward f32 + terminated u8 + truncated u8 per step",
           "k_rollout_mask<plain, episodes>", (1 << 18) * T0, ms, 7.0 + 16.0 * 4 / T0, 16.0 * 4 + 7.0,
           {"unit": "env-steps", "episodes_finished": int(episode.sum().item())})
    del eng, xy, pid, actions, reward, term, trunc, steps, episode
    torch.cuda.empty_cache()

    # ---- config 4: successor expansion of a 10M-state frontier ------------------------------
    eng = Engine([data.puzzle_text("level2/Pull Dont Push")], device=dev)
    lo, hi = distributed.shard_range(10_000_000, rank, world)
    n = hi - lo
    xy = eng.new_state(n)
    eng.reset(xy, 0)
    gen = torch.Generator(device=dev).manual_seed(1236 + rank)
    walk = torch.randint(0, 4, (64, n), dtype=torch.uint8, device=dev, generator=gen)
    eng.rollout(xy, 0, walk)  # frontier = 64-step seeded random walks
    del walk
    xy_out = torch.empty((2, eng.n_max, 4 * n), dtype=torch.int32, device=dev)
    moved = torch.empty(4 * n, dtype=torch.int32, device=dev)
    goal = torch.empty(4 * n, dtype=torch.uint8, device=dev)
    sampler.mark_begin()
    ms = _timed(lambda: eng.expand(xy, 0, xy_out=xy_out, moved=moved, goal=goal), reps, 2, dev, distributed)
    sampler.mark_end()
    record("cfg4_expand", "cfg4: level2/Pull Dont Push, 10M-state frontier -> 40M successors + moved mask + goal flag",
           "k_expand_uniform<8>", 40_000_000, ms, 10.0 * eng.n_max + 5.0, 10.0 * eng.n_max + 5.0,
           {"unit": "successors"})
    del xy_out, moved, goal, xy, eng
    torch.cuda.empty_cache()

    # ---- config 5: step + RGB observation, 2^16 environments --------------------------------
    eng = Engine([data.puzzle_text(PUZZLE_KEY)], device=dev)
    lo, hi = distributed.shard_range(1 << 16, rank, world)
    n = hi - lo
    xy = eng.new_state(n)
    eng.reset(xy, 0)
    act = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
    for dtype, size in (("float32", 4), ("uint8", 1)):
        obs = torch.empty((n, 100, 160, 3), dtype=torch.float32 if size == 4 else torch.uint8, device=dev)
        step_out = eng.step_render(xy, 0, act, dtype=dtype, out=obs)[1]
        sampler.mark_begin()
        ms = _timed(lambda: eng.step_render(xy, 0, act, dtype=dtype, out=obs, step_out=step_out),
                    reps, 2, dev, distributed)
        sampler.mark_end()
        record(f"cfg5_step_render_{'f32' if size == 4 else 'u8'}",
               f"cfg5: {PUZZLE_KEY}, 2^16 envs, pw_step_render, {dtype} observation (100,160,3)",
               "k_step + k_render", 1 << 16, ms, 100 * 160 * 3 * size + 54.0, 100 * 160 * 3 * size + 54.0,
               {"unit": "frames"})
        del obs
    del eng, xy
    torch.cuda.empty_cache()
    return out


def run_latency(dev):
    """Per-call latency of the per-step surfaces (rank 0, 1 GPU): the N=1 gym drop-in's step()
    (action in, step + render, observation + reward + flags + state back on the host) beside the
    CPU port of the reference's PushWorldEnv.step, and one vector-environment step of 2^18
    environments replayed as a CUDA graph."""
    import random

    import torch

    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import pushworld_oracle as po
    from pushworld_b200 import data
    from pushworld_b200.gym_env import PushWorldEnv
    from pushworld_b200.vector_env import PushWorldVectorEnv

    root = data.materialize()
    path = os.path.join(root, PUZZLE_KEY + ".pwp")
    env = PushWorldEnv(path)
    env.reset()
    rng = random.Random(7)
    for _ in range(20):
        env.step(rng.randrange(4))
    n_steps = 300
    t0 = time.perf_counter()
    for _ in range(n_steps):
        env.step(rng.randrange(4))
    dropin_us = (time.perf_counter() - t0) / n_steps * 1e6
    # the reference's step() restated on the CPU (oracle/pushworld_oracle.py: OracleEnv), same puzzle
    cpu_env = po.OracleEnv([po.OraclePuzzle(data.puzzle_text(PUZZLE_KEY))])
    cpu_env.reset()
    t0 = time.perf_counter()
    n_cpu = 0
    while time.perf_counter() - t0 < 2.0:
        cpu_env.step(rng.randrange(4))
        n_cpu += 1
    cpu_us = (time.perf_counter() - t0) / n_cpu * 1e6
    # vector step, 2^18 environments, no observation: one CUDA-graph replay per step
    venv = PushWorldVectorEnv([data.puzzle_text(PUZZLE_KEY)], N_ENVS, puzzle_ids=0, device=dev)
    venv.reset()
    step = venv.capture_step()
    step.actions.copy_(torch.randint(0, 4, (N_ENVS,), dtype=torch.uint8, device=dev))
    for _ in range(20):
        step.replay()
    torch.cuda.synchronize()
    n_rep = 500
    t0 = time.perf_counter()
    for _ in range(n_rep):
        step.replay()
    torch.cuda.synchronize()
    vec_us = (time.perf_counter() - t0) / n_rep * 1e6
    return {
        "dropin_gym_step_us": dropin_us,
        "dropin_gym_steps_per_s": 1e6 / dropin_us,
        "cpu_port_gym_step_us": cpu_us,
        "cpu_port_gym_steps_per_s": 1e6 / cpu_us,
        "dropin_vs_cpu_port": cpu_us / dropin_us,
        "vector_step_graph_replay_us": vec_us,
        "vector_step_env_steps_per_s": N_ENVS / (vec_us * 1e-6),
        "note": "N=1 drop-in: pushworld_b200.gym_env.PushWorldEnv.step on level1/Road Block, observation "
        "(100,160,3) float32 + reward + flags + state read back through pinned buffers, one CUDA-graph "
        "replay + one stream synchronisation per call; vector step: PushWorldVectorEnv.capture_step().replay() "
        "for 2^18 environments (wall clock per call incl. the launch, device-synchronised at the end)",
    }


def run_ours(args) -> None:
    import torch

    from pushworld_b200 import data, distributed
    from pushworld_b200.engine import record_to_outputs
    from pushworld_b200.vector_env import PushWorldVectorEnv, RolloutOutput

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    # The only collective is a 32-byte all-reduce per rollout, overlapped with the next rollout, whose 128 CTAs
    # need 128 of the 148 SMs at once: with NCCL's default channel count its (spinning) CTAs take more than the 20
    # free SMs and push part of the rollout into a second wave (+0.035 ms per 0.31 ms step at N = 8, measured).
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
    rank, local_rank, world = distributed.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # pinned host buffers (the e2e legs) are allocated after this: local to the GPU's NUMA node
    numa = distributed.bind_to_gpu_numa_node(local_rank)

    env = PushWorldVectorEnv([data.puzzle_text(PUZZLE_KEY)], N_ENVS, puzzle_ids=0, device=dev)
    n_obj = env.engine.n_max
    gen = torch.Generator(device=dev).manual_seed(ACTION_SEED + rank)
    actions = torch.randint(0, 4, (T_STEPS, N_ENVS), dtype=torch.uint8, device=dev, generator=gen)
    out = RolloutOutput(
        reward=torch.empty((T_STEPS, N_ENVS), dtype=torch.float32, device=dev),
        terminated=torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8, device=dev),
        counters=torch.zeros(4, dtype=torch.int64, device=dev),
    )
    # The path's single collective: ONE all-reduce (sum) of the int64[4] counter vector PER ROLLOUT.
    # It runs on a side stream and overlaps the next rollouts: the counter vectors are a ring of kRing
    # buffers, so the launching stream only waits for the all-reduce of kRing rollouts ago (with two
    # buffers an all-reduce that NCCL could not schedule next to the rollout's 128 resident CTAs held
    # the next-but-one rollout back: 0.05 ms per step at N = 8).
    kRing = 4
    # diagnostic only (PW_BENCH_NO_ALLREDUCE=1): one all-reduce at the end instead of one per rollout
    skip_allreduce = os.environ.get("PW_BENCH_NO_ALLREDUCE") == "1"
    side = torch.cuda.Stream(dev)
    counters = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(kRing)]
    reduced = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(kRing)]
    totals = torch.zeros(4, dtype=torch.int64, device=dev)
    ev_done = [torch.cuda.Event() for _ in range(kRing)]
    ev_free = [torch.cuda.Event() for _ in range(kRing)]

    def one_step(k, ev0=None, ev1=None):
        b = k % kRing
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev_free[b])  # the all-reduce of rollout k - kRing has consumed counters[b]
        counters[b].zero_()
        out.counters = counters[b]
        env.reset()
        if ev0 is not None:
            ev0.record()
        env.rollout(actions, out=out)
        if ev1 is not None:
            ev1.record()
        ev_done[b].record(cur)
        with torch.cuda.stream(side):
            side.wait_event(ev_done[b])
            reduced[b].copy_(counters[b])
            if not skip_allreduce:
                distributed.allreduce_counters(reduced[b])  # NCCL, int64[4]
            totals.add_(reduced[b])
            ev_free[b].record(side)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        one_step(k)
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs ------------------------------------------
    totals.zero_()
    kernel_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    distributed.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    start.record()
    host_t0 = time.perf_counter()
    for k in range(args.steps):
        one_step(k, *kernel_events[k])
    host_issue_ms = (time.perf_counter() - host_t0) * 1e3 / max(args.steps, 1)  # launch work of one step, host side
    torch.cuda.current_stream(dev).wait_stream(side)  # the last all-reduce is part of the job
    if skip_allreduce:
        distributed.allreduce_counters(totals)
    stop.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    distributed.barrier()
    elapsed_ms = distributed.max_over_ranks(start.elapsed_time(stop), dev)
    host_issue_ms = distributed.max_over_ranks(host_issue_ms, dev)
    total_env_steps = int(totals[0].item())
    assert total_env_steps == world * args.steps * N_ENVS * T_STEPS, total_env_steps
    value = total_env_steps / (elapsed_ms * 1e-3)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kernel_events)
    kernel_ms_slowest = distributed.max_over_ranks(kernel_ms, dev)  # the GPU that sets the max-over-ranks time
    out.counters = counters[0]

    # ---- end to end: pinned host buffers in, pinned host buffers out ---------------------
    # (1) the headline: actions in (1 B / env-step), packed per-step records out (1 B / env-step)
    h_act = torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8).pin_memory()
    h_act.copy_(actions)
    h_rec = torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8).pin_memory()
    e2e_steps = min(args.steps, 40)
    chunks = (T_STEPS + 63) // 64

    def e2e_leg(step_fn):
        for _ in range(min(args.warmup, 3)):
            step_fn()
        distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.mark_begin()
        a.record()
        for _ in range(e2e_steps):
            step_fn()
        b.record()
        torch.cuda.synchronize()
        sampler.mark_end()
        ms = distributed.max_over_ranks(a.elapsed_time(b), dev)
        return ms, world * e2e_steps * N_ENVS * T_STEPS / (ms * 1e-3)

    def e2e_packed_step():
        env.reset()
        env.rollout_host_packed(h_act, h_rec, chunk_steps=64)

    e2e_ms, e2e_value = e2e_leg(e2e_packed_step)
    # the records must expand to exactly the float32 rewards / flags of the device-resident run
    r_last, t_last, _, _ = record_to_outputs(h_rec[-1])
    assert torch.equal(r_last, out.reward[-1].cpu()) and torch.equal(t_last, out.terminated[-1].cpu())
    r_mid, t_mid, _, _ = record_to_outputs(h_rec[T_STEPS // 2, : 1 << 14])
    assert torch.equal(r_mid, out.reward[T_STEPS // 2, : 1 << 14].cpu())
    del h_rec

    # (2) the same with the reference's own output types: float32 reward + uint8 terminated out (5 B)
    h_rew = torch.empty((T_STEPS, N_ENVS), dtype=torch.float32).pin_memory()
    h_term = torch.empty((T_STEPS, N_ENVS), dtype=torch.uint8).pin_memory()

    def e2e_f32_step():
        env.reset()
        env.rollout_host(h_act, h_rew, h_term, chunk_steps=64)

    e2e_f32_ms, e2e_f32_value = e2e_leg(e2e_f32_step)
    assert torch.equal(h_rew[-1], out.reward[-1].cpu()) and torch.equal(h_term[-1], out.terminated[-1].cpu())
    del h_rew, h_term

    # (3) summaries only: host actions in, per-environment returns out
    h_ret = torch.empty(N_ENVS, dtype=torch.float32).pin_memory()
    h_nterm = torch.empty(N_ENVS, dtype=torch.int32).pin_memory()

    def e2e_returns_step():
        env.reset()
        env.rollout_host_returns(h_act, h_ret, h_nterm, chunk_steps=64)

    e2e_ret_ms, e2e_ret_value = e2e_leg(e2e_returns_step)
    assert torch.equal(h_nterm.to(torch.int64), out.terminated.sum(dim=0, dtype=torch.int64).cpu())
    del h_act, env, actions, out
    torch.cuda.empty_cache()

    peak, peak_src = _peak()
    configs = run_configs(rank, world, dev, peak, sampler) if not args.no_configs else {}
    latency = run_latency(dev) if (rank == 0 and not args.no_configs) else None
    distributed.barrier()
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (pw_rollout) -----------------------------------
    units = N_ENVS * T_STEPS
    # bytes a T-fused rollout must move per env-step: action 1 + reward 4 + terminated 1, plus the
    # state read once and written once per launch (16 n / T).  SURVEY 8(d)'s 16n+6 is the T=1 figure.
    fused_bpu = 6.0 + 16.0 * n_obj / T_STEPS
    step_bpu = 16.0 * n_obj + 6.0
    achieved = units * fused_bpu / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": ROLLOUT_KERNEL,
        "kernel_ms_slowest_rank": kernel_ms_slowest,
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
        roofline["traffic_source"] = traffic.get("source")
        # secondary bound: what limits the kernel when it is not HBM
        roofline["secondary_bound"] = traffic.get("secondary_bound")

    cpu_rate, cpu_threads, cpu_sample, cpu_kind = cpu_throughput(12.0)
    port_rate = cpu_throughput(4.0, kind="port")[0] if cpu_kind == "reference" else cpu_rate
    py_rate = python_port_throughput(3.0)

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps,
        "host_issue_ms_per_step": host_issue_ms,  # max over ranks: the launch work of a step on the host
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32",
        "data": "synthetic",
        "config": _workload_config(world),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": cpu_kind,
                         "sample": cpu_sample,
                         "c_port_all_cores": port_rate,
                         "python_port_single_core": py_rate},
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": T_STEPS * N_ENVS,
            "d2h_bytes_per_step": T_STEPS * N_ENVS,
            "ms_per_step": e2e_ms / e2e_steps,
            "api": "PushWorldVectorEnv.rollout_host_packed (pinned host buffers, 3-stream pipeline): "
            "uint8 actions in, one packed record byte per env-step out (delta_goals | truncated | "
            "terminated; engine.record_to_outputs expands it to the reference's float reward + flags, "
            "checked bit-exact against the device-resident run)",
            "numa": numa,
        },
        "e2e_f32": {
            "value": e2e_f32_value,
            "unit": UNIT,
            "h2d_bytes_per_step": T_STEPS * N_ENVS,
            "d2h_bytes_per_step": T_STEPS * N_ENVS * 5,
            "ms_per_step": e2e_f32_ms / e2e_steps,
            "api": "PushWorldVectorEnv.rollout_host (float32 reward + uint8 terminated out: the round-1 e2e)",
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
        "configs": configs,
        "per_step_latency": latency,
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
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the BASELINE config 3/4/5 sub-records (headline + e2e only)")
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

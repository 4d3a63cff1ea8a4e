#!/usr/bin/env python3
"""A/B measurements of the general-rule rollout kernels on BASELINE config 3 (1 GPU):

    python tools/occ_probe.py [--steps 256] [--out gpurun_out/occ_probe.json]

* k_rollout_occ (occupancy reject) with the push closure per thread and per warp (PW_OCC_WARP_CLOSURE=1),
* the scan-every-step kernels it replaces (PW_NO_OCC=1),
* the warp-per-environment layout of BASELINE.json's north star (PW_WARP_PER_ENV=1),
each on the grouped and on the `e mod 223` interleaved batch, checked against each other.
"""

import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from pushworld_b200 import data  # noqa: E402
from pushworld_b200.engine import Engine  # noqa: E402


def time_ms(fn, warmup=2, reps=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    keys = data.benchmark_keys()
    eng = Engine([data.puzzle_text(k) for k in keys], device=dev)
    n, T = args.envs, args.steps
    ids = torch.arange(n, dtype=torch.int64) % len(keys)
    layouts = {"grouped": torch.sort(ids, stable=True).values.to(torch.int32).to(dev),
               "interleaved": ids.to(torch.int32).to(dev)}
    gen = torch.Generator(device=dev).manual_seed(1235)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device=dev, generator=gen)
    reward = torch.empty((T, n), dtype=torch.float32, device=dev)
    term = torch.empty((T, n), dtype=torch.uint8, device=dev)
    xy = eng.new_state(n)
    results = []
    variants = [("bounding-box mask prefilter, lock-step (k_rollout_mask)", {}),
                ("mask prefilter, closure work compacted across the CTA, two barriers per step", {"PW_MASK_CTA": "1"})]
    variants += [("occupancy reject, run-ahead lanes, batched closure (k_rollout_occ_ra) batch=%d" % b,
                  {"PW_OCC_MODE": "runahead", "PW_OCC_BATCH": str(b)}) for b in (8,)]
    variants += [("occupancy reject, lock-step, per-thread closure (k_rollout_occ)", {"PW_OCC_MODE": "lockstep"}),
                ("occupancy reject, lock-step, warp-per-environment closure",
                 {"PW_OCC_MODE": "lockstep", "PW_OCC_WARP_CLOSURE": "1"}),
                ("scan every step (round 1)", {"PW_NO_OCC": "1"}),
                ("warp per environment for every step", {"PW_WARP_PER_ENV": "1"})]
    for layout, pid in layouts.items():
        ref = None
        for name, env in variants:
            for k in ("PW_OCC_WARP_CLOSURE", "PW_OCC_MODE", "PW_OCC_BATCH", "PW_NO_OCC", "PW_WARP_PER_ENV", "PW_MASK_CTA"):
                os.environ.pop(k, None)
            os.environ.update(env)

            def run():
                eng.reset(xy, pid)
                eng.rollout(xy, pid, actions, reward=reward, terminated=term)

            ms = time_ms(run)
            digest = (int(reward.view(torch.int32).to(torch.int64).sum()), int(term.sum(dtype=torch.int64)),
                      int(xy.to(torch.int64).sum()))
            if ref is None:
                ref = (reward.clone(), term.clone(), xy.clone())
            same = torch.equal(reward, ref[0]) and torch.equal(term, ref[1]) and torch.equal(xy, ref[2])
            rec = {"layout": layout, "variant": name, "kernel_ms": ms, "env_steps_per_s": n * T / (ms * 1e-3),
                   "equal_to_first_variant": bool(same), "digest": digest}
            print(json.dumps(rec), flush=True)
            results.append(rec)
    for k in ("PW_OCC_WARP_CLOSURE", "PW_OCC_MODE", "PW_OCC_BATCH", "PW_NO_OCC", "PW_WARP_PER_ENV", "PW_MASK_CTA"):
        os.environ.pop(k, None)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()

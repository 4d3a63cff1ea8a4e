import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from pushworld_b200 import data
from pushworld_b200.vector_env import PushWorldVectorEnv, RolloutOutput
dev = torch.device("cuda", 0)
N, T = 1 << 18, 1024
env = PushWorldVectorEnv([data.puzzle_text("level1/Road Block")], N, puzzle_ids=0, device=dev)
actions = torch.randint(0, 4, (T, N), dtype=torch.uint8, device=dev)
out = RolloutOutput(reward=torch.empty((T, N), dtype=torch.float32, device=dev), terminated=torch.empty((T, N), dtype=torch.uint8, device=dev), counters=torch.zeros(4, dtype=torch.int64, device=dev))
small_T = actions[:8].contiguous()
small = RolloutOutput(reward=torch.empty((8, N), dtype=torch.float32, device=dev), terminated=torch.empty((8, N), dtype=torch.uint8, device=dev), counters=out.counters)
def timeit(name, fn, reps=300):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host issue {1e6*(t1-t0)/reps:.1f} us/call, total {1e6*(t2-t0)/reps:.1f} us/call", flush=True)
timeit("env.reset", env.reset)
timeit("env.rollout T=8", lambda: env.rollout(small_T, out=small))
timeit("counters.zero_", out.counters.zero_)
ev = torch.cuda.Event(enable_timing=True)
timeit("event.record", ev.record)
e2 = torch.cuda.Event()
cur = torch.cuda.current_stream(dev)
timeit("wait_event", lambda: cur.wait_event(e2))
eng = env.engine
xy = env.xy if hasattr(env, "xy") else None
print([a for a in dir(env) if not a.startswith("__")][:60])

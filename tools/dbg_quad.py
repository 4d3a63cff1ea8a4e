import os, sys, torch
sys.path.insert(0, "/root/repo")
from pushworld_b200 import data
from pushworld_b200.engine import Engine, record_to_outputs
text = data.puzzle_text("level1/Road Block")
eng = Engine([text])
n = int(os.environ.get("DBG_N", "256"))
for T in (16, 32, 13, 333):
    gen = torch.Generator(device="cuda").manual_seed(41)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    xy0 = eng.new_state(n); eng.reset(xy0, 0)
    outs = {}
    for name, envs in (("quad", {}), ("pairs", {"PW_NO_QUAD": "1"})):
        for k in ("PW_NO_QUAD",): os.environ.pop(k, None)
        os.environ.update(envs)
        xy = xy0.clone()
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
        term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        eng.rollout(xy, 0, actions, reward=reward, terminated=term, counters=counters)
        xy2 = xy0.clone(); c2 = torch.zeros(4, dtype=torch.int64, device="cuda")
        record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        eng.rollout_packed(xy2, 0, actions, record, counters=c2)
        outs[name] = (xy, reward, term, counters, xy2, record, c2)
    names = ["xy", "reward", "term", "counters", "xy_p", "record", "counters_p"]
    for nm, a, b in zip(names, outs["quad"], outs["pairs"]):
        if not torch.equal(a, b):
            bad = (a != b).nonzero()
            print("T", T, nm, "mismatches", bad.shape[0], "first", bad[0].tolist(), a[tuple(bad[0])].item(), b[tuple(bad[0])].item())
            if nm == "counters": print(a.tolist(), b.tolist())
        else:
            print("T", T, nm, "ok")

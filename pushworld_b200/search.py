"""Batched breadth-first search on the GPU: the "next" step after successor expansion
(SURVEY.md 8f-3).  The reference's planner pops one node at a time and calls ``getNextState`` four
times per node (cpp/include/search/best_first_search.h:72-92); here a whole BFS layer is expanded by
one ``pw_expand_indexed`` launch and de-duplicated exactly by one ``pw_visited_insert`` launch.
PyTorch only compacts index lists (``nonzero``) and holds the arena.

This is uninformed search (no RGD heuristic, which stays out of scope): it returns a SHORTEST plan
and is meant for small puzzles and as an end-to-end exercise of the expansion path.
"""

from __future__ import annotations

from typing import Optional, Union

import torch

from .engine import Engine

ACTION_CHARS = "LRUD"  # cpp/include/pushworld_puzzle.h:71


def bfs_solve(puzzle: Union[str, Engine], index: int = 0, max_states: int = 1 << 24,
              device=None, return_stats: bool = False):
    """Shortest plan (string over ``LRUD``) that reaches the goal from the initial state, or None
    if the reachable states (up to ``max_states`` generated successors) contain no goal state."""
    engine = puzzle if isinstance(puzzle, Engine) else Engine([puzzle], device=device)
    dev = engine.device
    cap = int(max_states)
    arena = torch.zeros((2, engine.n_max, cap), dtype=torch.int32, device=dev)
    table_cap = 1 << max(10, (2 * cap - 1).bit_length())
    table = torch.zeros(table_cap, dtype=torch.int64, device=dev)
    parent = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    via = torch.zeros(cap, dtype=torch.uint8, device=dev)

    # state 0 = initial state
    init = engine.new_state(1)
    engine.reset(init, index)
    arena[:, :, 0:1] = init
    is_new = torch.empty(1, dtype=torch.uint8, device=dev)
    engine.visited_insert(arena, index, 0, 1, table, is_new)
    _, is_goal = engine.goals(init, index)
    stats = {"layers": 0, "generated": 1, "unique": 1}
    if bool(is_goal.item()):
        return ("", stats) if return_stats else ""

    frontier = torch.zeros(1, dtype=torch.int64, device=dev)
    used = 1
    goal_id = -1
    while frontier.numel() > 0:
        f = frontier.numel()
        if used + 4 * f > cap:
            break  # arena exhausted
        moved = torch.empty(4 * f, dtype=torch.int32, device=dev)
        goal = torch.empty(4 * f, dtype=torch.uint8, device=dev)
        fresh = torch.empty(4 * f, dtype=torch.uint8, device=dev)
        engine.expand_indexed(arena, index, frontier, used, moved, goal)
        # a successor that did not move equals its parent: never new
        engine.visited_insert(arena, index, used, 4 * f, table, fresh, valid=(moved != 0).to(torch.uint8))
        new_local = torch.nonzero(fresh, as_tuple=False).flatten()
        new_ids = new_local + used
        parent[new_ids] = frontier[new_local % f]
        via[new_ids] = (new_local // f).to(torch.uint8)
        stats["layers"] += 1
        stats["generated"] += 4 * f
        stats["unique"] += int(new_ids.numel())
        hits = new_ids[goal[new_local] != 0]
        if hits.numel() > 0:
            goal_id = int(hits.min().item())
            break
        used += 4 * f
        frontier = new_ids
    if engine.poll_status()[1]:
        raise RuntimeError("visited table overflow")
    if goal_id < 0:
        return (None, stats) if return_stats else None
    plan = []
    node = goal_id
    while node != 0:
        plan.append(ACTION_CHARS[int(via[node].item())])
        node = int(parent[node].item())
    plan = "".join(reversed(plan))
    return (plan, stats) if return_stats else plan

// ref_wrap.cc -- C entry points around the REFERENCE's own C++ simulator and novelty heuristic,
// compiled from the sources where they lie (/root/reference/cpp/src/pushworld_puzzle.cc,
// /root/reference/cpp/src/heuristics/novelty.cc; recipe: oracle/Makefile, target _ref) into
// oracle/_ref/libpw_ref.so.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
// bench.py's CPU-baseline / --impl reference legs may load it.  Nothing here restates the
// reference's algorithm -- it only converts between the reference's types
// (State = vector<Position2D>, Position2D = x * POSITION_LIMIT + y, pushworld_puzzle.h:32-40) and
// plain int arrays, and loops over environments.
//
// Object ORDER is the C++ one (ascending std::map, pushworld_puzzle.cc:274-315): "a", the movable
// of every goal id in ascending order, then the remaining movables ascending.  The python
// reference -- the parity target -- orders objects differently (puzzle.py:178-179,235-237); callers
// permute (oracle/ref_oracle.py).
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "heuristics/novelty.h"
#include "pushworld_puzzle.h"

#ifdef _OPENMP
#include <omp.h>
#endif

using pushworld::POSITION_LIMIT;
using pushworld::PushWorldPuzzle;
using pushworld::State;

static thread_local std::string g_error;

static inline int pack(int x, int y) { return x * POSITION_LIMIT + y; }
static inline void unpack(int p, int* x, int* y) {
  // positions are non-negative inside the grid; floor division keeps x right for y in [0, LIMIT)
  *x = p / POSITION_LIMIT;
  *y = p % POSITION_LIMIT;
}

static int achieved(const PushWorldPuzzle& puzzle, const State& s) {  // puzzle.py:396-407
  const auto& goal = puzzle.getGoal();
  int count = 0;
  for (size_t k = 0; k < goal.size(); ++k) count += s[k + 1] == goal[k];
  return count;
}

extern "C" {

const char* pwr_error(void) { return g_error.c_str(); }

void* pwr_load(const char* path) {
  try {
    return new PushWorldPuzzle(std::string(path));
  } catch (const std::exception& ex) {
    g_error = ex.what();
    return nullptr;
  }
}

void pwr_free(void* h) { delete static_cast<PushWorldPuzzle*>(h); }
int pwr_num_objects(void* h) { return (int)static_cast<PushWorldPuzzle*>(h)->getInitialState().size(); }
int pwr_num_goals(void* h) { return (int)static_cast<PushWorldPuzzle*>(h)->getGoal().size(); }

void pwr_initial_state(void* h, int* xy) {
  const State& s = static_cast<PushWorldPuzzle*>(h)->getInitialState();
  for (size_t k = 0; k < s.size(); ++k) unpack(s[k], &xy[2 * k], &xy[2 * k + 1]);
}

void pwr_goal(void* h, int* xy) {
  const auto& g = static_cast<PushWorldPuzzle*>(h)->getGoal();
  for (size_t k = 0; k < g.size(); ++k) unpack(g[k], &xy[2 * k], &xy[2 * k + 1]);
}

// getNextState (pushworld_puzzle.cc:386-460): xy updated in place, returns the bit mask of
// RelativeState.moved_object_indices.
uint64_t pwr_step(void* h, int* xy, int action) {
  const PushWorldPuzzle& puzzle = *static_cast<PushWorldPuzzle*>(h);
  const size_t n = puzzle.getInitialState().size();
  State s(n);
  for (size_t k = 0; k < n; ++k) s[k] = pack(xy[2 * k], xy[2 * k + 1]);
  const auto next = puzzle.getNextState(s, action);
  for (size_t k = 0; k < n; ++k) unpack(next.state[k], &xy[2 * k], &xy[2 * k + 1]);
  uint64_t mask = 0;
  for (int i : next.moved_object_indices) mask |= 1ull << i;
  return mask;
}

int pwr_satisfies_goal(void* h, const int* xy) {
  const PushWorldPuzzle& puzzle = *static_cast<PushWorldPuzzle*>(h);
  const size_t n = puzzle.getInitialState().size();
  State s(n);
  for (size_t k = 0; k < n; ++k) s[k] = pack(xy[2 * k], xy[2 * k + 1]);
  return puzzle.satisfiesGoal(s) ? 1 : 0;
}

// T steps of n_envs environments, SoA state x[j * stride + e] in the C++ object order, with the
// reward / terminated logic of PushWorldEnv.step (gym_env.py:201-223).  getNextState is not
// re-entrant (mutable scratch members, pushworld_puzzle.h:178-180): every thread steps its own COPY
// of the puzzle.  Returns the number of env-steps executed.
int64_t pwr_rollout(void* h, int64_t n_envs, int n_steps, int32_t* x, int32_t* y, int64_t stride,
                    const uint8_t* actions, float* reward, uint8_t* terminated, int threads) {
  const PushWorldPuzzle& shared = *static_cast<PushWorldPuzzle*>(h);
  const int n = (int)shared.getInitialState().size();
  const int g = (int)shared.getGoal().size();
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
  {
    const PushWorldPuzzle puzzle(shared);  // private scratch
    State s(n);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t e = 0; e < n_envs; ++e) {
      for (int k = 0; k < n; ++k) s[k] = pack(x[k * stride + e], y[k * stride + e]);
      int prev = achieved(puzzle, s);
      for (int t = 0; t < n_steps; ++t) {
        const int64_t idx = (int64_t)t * n_envs + e;
        s = puzzle.getNextState(s, actions[idx]).state;
        const int ach = achieved(puzzle, s);
        const bool term = ach == g;
        if (reward) reward[idx] = term ? 10.0f : (float)((double)(ach - prev) - 0.01);
        if (terminated) terminated[idx] = (uint8_t)term;
        prev = ach;
      }
      for (int k = 0; k < n; ++k) {
        int px, py;
        unpack(s[k], &px, &py);
        x[k * stride + e] = px;
        y[k * stride + e] = py;
      }
    }
  }
  (void)threads;
  return n_envs * (int64_t)n_steps;
}

// NoveltyHeuristic (cpp/src/heuristics/novelty.cc:30-77)
void* pwr_novelty_create(int state_size) { return new pushworld::heuristic::NoveltyHeuristic(state_size); }
void pwr_novelty_free(void* h) { delete static_cast<pushworld::heuristic::NoveltyHeuristic*>(h); }
int pwr_novelty(void* h, const int* xy, int n, uint64_t moved_mask) {
  pushworld::RelativeState rs;
  rs.state.resize(n);
  for (int k = 0; k < n; ++k) rs.state[k] = pack(xy[2 * k], xy[2 * k + 1]);
  for (int k = 0; k < n; ++k)
    if ((moved_mask >> k) & 1ull) rs.moved_object_indices.push_back(k);
  return (int)static_cast<pushworld::heuristic::NoveltyHeuristic*>(h)->estimate_cost_to_goal(rs);
}

}  // extern "C"

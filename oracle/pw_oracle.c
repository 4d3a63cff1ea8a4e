/*
 * pw_oracle.c -- plain-C CPU restatement of the PushWorld transition, TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  It is the checker and the CPU baseline, never the product path.
 *
 * Unlike the CUDA kernels (row bitmasks, relative geometry evaluated on the fly) this file
 * follows the REFERENCE's own algorithm: collision relations are precomputed into membership
 * tables at load time and a step is a depth-first walk over table look-ups.
 *   - static_tab[a][k]     <-> _agent_collision_map / _wall_collision_map
 *                              (python3/src/pushworld/puzzle.py:262-293, builder :522-564)
 *   - dynamic_tab[a][i][j] <-> _movable_collision_map (puzzle.py:266-308, builder :567-593)
 *   - pwo_step             <-> PushWorldPuzzle.get_next_state (puzzle.py:348-394), same LIFO
 *                              frontier as the reference; moved mask as in
 *                              cpp/src/pushworld_puzzle.cc:439-452
 *   - pwo_achieved         <-> count_achieved_goals (puzzle.py:396-407)
 *   - pwo_rollout          <-> repeated PushWorldEnv.step reward logic (gym_env.py:201-223)
 * The tables are dense bitmaps (grid positions / relative offsets) instead of hash sets.
 * Parsing is done by oracle/pushworld_oracle.py, which hands over plain cell lists.
 *
 * Parity: PINNED -- tests/test_oracle_golden.py checks this file against the vectors the real
 * reference produced for all 223 benchmark puzzles (tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PWO_MAX_OBJ 64

typedef struct {
  int W, H, n, g;
  int *ncells;            /* [n] */
  int **cells;            /* [n][2*ncells] relative (x, y) */
  int bw[PWO_MAX_OBJ], bh[PWO_MAX_OBJ];
  int init[2 * PWO_MAX_OBJ], goal[2 * PWO_MAX_OBJ];
  uint8_t *wall, *awall;  /* [W*H] occupancy grids */
  /* static_tab[a][k]: bitmap over positions (x, y) in [0,W) x [0,H) */
  uint8_t *static_tab;    /* [4][n][W*H] */
  /* dynamic_tab[a][i][j]: bitmap over rel = pos_i - pos_j, each coordinate in [-R, R] */
  int R, D;               /* D = 2R+1 */
  uint8_t *dynamic_tab;   /* [4][n][n][D*D] */
} PwoPuzzle;

static const int DX[4] = {-1, 1, 0, 0}; /* puzzle.py:43-50 */
static const int DY[4] = {0, 0, -1, 1};

static int grid_at(const uint8_t *grid, int W, int H, int x, int y) {
  return x >= 0 && y >= 0 && x < W && y < H && grid[y * W + x];
}

/* does object k placed at (px, py) overlap the static grid? */
static int overlaps_static(const PwoPuzzle *p, int k, int px, int py, const uint8_t *grid) {
  for (int c = 0; c < p->ncells[k]; ++c)
    if (grid_at(grid, p->W, p->H, px + p->cells[k][2 * c], py + p->cells[k][2 * c + 1])) return 1;
  return 0;
}

/* does object i at offset (ox, oy) relative to object j's frame overlap object j?
 * occ_j: bh_j x bw_j occupancy bitmap of object j. */
static int overlaps_pair(const PwoPuzzle *p, int i, int j, const uint8_t *occ_j, int ox, int oy) {
  for (int a = 0; a < p->ncells[i]; ++a) {
    const int x = p->cells[i][2 * a] + ox, y = p->cells[i][2 * a + 1] + oy;
    if (x >= 0 && y >= 0 && x < p->bw[j] && y < p->bh[j] && occ_j[y * p->bw[j] + x]) return 1;
  }
  return 0;
}

void pwo_free(PwoPuzzle *p) {
  if (!p) return;
  if (p->cells)
    for (int k = 0; k < p->n; ++k) free(p->cells[k]);
  free(p->cells);
  free(p->ncells);
  free(p->wall);
  free(p->awall);
  free(p->static_tab);
  free(p->dynamic_tab);
  free(p);
}

/* cells_xy: concatenated relative (x, y) pairs of all objects; walls given as absolute cells;
 * awall must already be aw UNION w (puzzle.py:273). */
PwoPuzzle *pwo_create(int W, int H, int n, int g, const int *ncells, const int *cells_xy,
                      const int *init_xy, const int *goal_xy, int n_wall, const int *wall_xy,
                      int n_awall, const int *awall_xy) {
  if (n < 1 || n > PWO_MAX_OBJ || g > n) return NULL;
  PwoPuzzle *p = (PwoPuzzle *)calloc(1, sizeof(PwoPuzzle));
  p->W = W; p->H = H; p->n = n; p->g = g;
  p->ncells = (int *)malloc(sizeof(int) * n);
  p->cells = (int **)calloc(n, sizeof(int *));
  int maxdim = W > H ? W : H;
  const int *src = cells_xy;
  for (int k = 0; k < n; ++k) {
    p->ncells[k] = ncells[k];
    p->cells[k] = (int *)malloc(sizeof(int) * 2 * ncells[k]);
    memcpy(p->cells[k], src, sizeof(int) * 2 * ncells[k]);
    src += 2 * ncells[k];
    int mx = 0, my = 0;
    for (int c = 0; c < ncells[k]; ++c) {
      if (p->cells[k][2 * c] > mx) mx = p->cells[k][2 * c];
      if (p->cells[k][2 * c + 1] > my) my = p->cells[k][2 * c + 1];
    }
    p->bw[k] = mx + 1; p->bh[k] = my + 1;
    if (p->bw[k] > maxdim) maxdim = p->bw[k];
    if (p->bh[k] > maxdim) maxdim = p->bh[k];
  }
  memcpy(p->init, init_xy, sizeof(int) * 2 * n);
  memcpy(p->goal, goal_xy, sizeof(int) * 2 * g);
  p->wall = (uint8_t *)calloc((size_t)W * H, 1);
  p->awall = (uint8_t *)calloc((size_t)W * H, 1);
  for (int c = 0; c < n_wall; ++c) p->wall[wall_xy[2 * c + 1] * W + wall_xy[2 * c]] = 1;
  for (int c = 0; c < n_awall; ++c) p->awall[awall_xy[2 * c + 1] * W + awall_xy[2 * c]] = 1;

  /* ---- static collision tables: puzzle.py:522-564 ---- */
  p->static_tab = (uint8_t *)calloc((size_t)4 * n * W * H, 1);
  for (int a = 0; a < 4; ++a)
    for (int k = 0; k < n; ++k) {
      const uint8_t *grid = k == 0 ? p->awall : p->wall; /* puzzle.py:272-293 */
      uint8_t *tab = p->static_tab + ((size_t)(a * n + k)) * W * H;
      /* positions with the bounding box inside the grid, puzzle.py:557-561 */
      for (int y = 0; y + p->bh[k] <= H; ++y)
        for (int x = 0; x + p->bw[k] <= W; ++x)
          tab[y * W + x] = overlaps_static(p, k, x + DX[a], y + DY[a], grid) &&
                           !overlaps_static(p, k, x, y, grid);
    }

  /* ---- dynamic collision tables: puzzle.py:567-593 ----
   * key = pos_pusher - pos_pushee (puzzle.py:370); the pusher, displaced by the action, must
   * overlap the pushee while the undisplaced pusher does not. */
  p->R = maxdim + 1;
  p->D = 2 * p->R + 1;
  const size_t dd = (size_t)p->D * p->D;
  p->dynamic_tab = (uint8_t *)calloc((size_t)4 * n * n * dd, 1);
  for (int j = 1; j < n; ++j) {
    uint8_t *occ = (uint8_t *)calloc((size_t)p->bw[j] * p->bh[j], 1);
    for (int c = 0; c < p->ncells[j]; ++c)
      occ[p->cells[j][2 * c + 1] * p->bw[j] + p->cells[j][2 * c]] = 1;
    for (int a = 0; a < 4; ++a)
      for (int i = 0; i < n; ++i) {
        if (i == j) continue;
        uint8_t *tab = p->dynamic_tab + (((size_t)(a * n + i)) * n + j) * dd;
        /* offsets at which the bounding boxes (grown by the displacement) can touch */
        for (int ry = -p->bh[i] - 1; ry <= p->bh[j] + 1; ++ry)
          for (int rx = -p->bw[i] - 1; rx <= p->bw[j] + 1; ++rx)
            tab[(ry + p->R) * p->D + (rx + p->R)] =
                overlaps_pair(p, i, j, occ, rx + DX[a], ry + DY[a]) &&
                !overlaps_pair(p, i, j, occ, rx, ry);
      }
    free(occ);
  }
  return p;
}

/* puzzle.py:348-394.  xy: [2n] interleaved (x0, y0, x1, y1, ...), updated in place.
 * Returns the moved-object bitmask (0 if the state is unchanged). */
uint64_t pwo_step(const PwoPuzzle *p, int *xy, int action) {
  const int n = p->n, W = p->W, H = p->H;
  const size_t wh = (size_t)W * H, dd = (size_t)p->D * p->D;
  if (action < 0 || action > 3) return 0;
  {
    const int x = xy[0], y = xy[1];
    if (x >= 0 && y >= 0 && x < W && y < H &&
        p->static_tab[((size_t)(action * n + 0)) * wh + (size_t)y * W + x])
      return 0; /* the agent cannot move, puzzle.py:353-354 */
  }
  int frontier[PWO_MAX_OBJ], nf = 0;
  uint64_t pushed = 1;
  frontier[nf++] = 0;
  while (nf) {
    const int i = frontier[--nf]; /* LIFO like list.pop(), puzzle.py:360 */
    for (int j = 1; j < n; ++j) {
      if ((pushed >> j) & 1) continue;
      const int rx = xy[2 * i] - xy[2 * j], ry = xy[2 * i + 1] - xy[2 * j + 1];
      if (rx < -p->R || rx > p->R || ry < -p->R || ry > p->R) continue;
      if (!p->dynamic_tab[(((size_t)(action * n + i)) * n + j) * dd + (size_t)(ry + p->R) * p->D +
                          (rx + p->R)])
        continue;
      const int x = xy[2 * j], y = xy[2 * j + 1];
      if (x >= 0 && y >= 0 && x < W && y < H &&
          p->static_tab[((size_t)(action * n + j)) * wh + (size_t)y * W + x])
        return 0; /* transitive stopping, puzzle.py:376-379 */
      pushed |= 1ull << j;
      frontier[nf++] = j;
    }
  }
  for (int k = 0; k < n; ++k)
    if ((pushed >> k) & 1) {
      xy[2 * k] += DX[action];
      xy[2 * k + 1] += DY[action];
    }
  return pushed;
}

int pwo_achieved(const PwoPuzzle *p, const int *xy) { /* puzzle.py:396-407 */
  int count = 0;
  for (int k = 0; k < p->g; ++k)
    count += xy[2 * (k + 1)] == p->goal[2 * k] && xy[2 * (k + 1) + 1] == p->goal[2 * k + 1];
  return count;
}

void pwo_initial_state(const PwoPuzzle *p, int *xy) { memcpy(xy, p->init, sizeof(int) * 2 * p->n); }
int pwo_num_objects(const PwoPuzzle *p) { return p->n; }

/* T steps of n_envs environments of ONE puzzle, SoA state x[j*stride+e] like the CUDA engine.
 * reward / terminated may be NULL.  Uses `threads` OpenMP threads (<=0: all).  Returns the
 * number of env-steps executed. */
int64_t pwo_rollout(const PwoPuzzle *p, int64_t n_envs, int n_steps, int32_t *x, int32_t *y,
                    int64_t stride, const uint8_t *actions, float *reward, uint8_t *terminated,
                    int threads) {
  const int n = p->n;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_envs; ++e) {
    int xy[2 * PWO_MAX_OBJ];
    for (int k = 0; k < n; ++k) {
      xy[2 * k] = x[k * stride + e];
      xy[2 * k + 1] = y[k * stride + e];
    }
    int prev = pwo_achieved(p, xy);
    for (int t = 0; t < n_steps; ++t) {
      const int64_t idx = (int64_t)t * n_envs + e;
      pwo_step(p, xy, actions[idx]);
      const int ach = pwo_achieved(p, xy);
      const int term = ach == p->g; /* puzzle.py:409-411 */
      if (reward) reward[idx] = term ? 10.0f : (float)((double)(ach - prev) - 0.01);
      if (terminated) terminated[idx] = (uint8_t)term;
      prev = ach;
    }
    for (int k = 0; k < n; ++k) {
      x[k * stride + e] = xy[2 * k];
      y[k * stride + e] = xy[2 * k + 1];
    }
  }
  return n_envs * (int64_t)n_steps;
}

/* Successor expansion of n_states states of ONE puzzle (cpp/include/search/best_first_search.h:
 * 76-92: 4 x getNextState + satisfiesGoal per node).  Successor s = a * n_states + e, SoA like
 * pw_expand: x_out[j * stride_out + s].  moved: uint32 mask as in cpp/src/pushworld_puzzle.cc:
 * 439-452; goal: uint8.  moved / goal may be NULL. */
void pwo_expand(const PwoPuzzle *p, int64_t n_states, const int32_t *x, const int32_t *y,
                int64_t stride, int32_t *x_out, int32_t *y_out, int64_t stride_out,
                uint32_t *moved, uint8_t *goal, int threads) {
  const int n = p->n;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_states; ++e) {
    for (int a = 0; a < 4; ++a) {
      int xy[2 * PWO_MAX_OBJ];
      for (int k = 0; k < n; ++k) {
        xy[2 * k] = x[k * stride + e];
        xy[2 * k + 1] = y[k * stride + e];
      }
      const uint64_t m = pwo_step(p, xy, a);
      const int64_t s = (int64_t)a * n_states + e;
      for (int k = 0; k < n; ++k) {
        x_out[k * stride_out + s] = xy[2 * k];
        y_out[k * stride_out + s] = xy[2 * k + 1];
      }
      if (moved) moved[s] = (uint32_t)m;
      if (goal) goal[s] = (uint8_t)(pwo_achieved(p, xy) == p->g);
    }
  }
}

int pwo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// pw_render_core.cuh -- per-cell closed form of PushWorldPuzzle.render (host/device code).
//
// The reference paints objects layer by layer (python3/src/pushworld/puzzle.py:450-467) and,
// per object cell, first the fill and then a border strip of `border_width` pixels towards each
// of the 8 neighbours that is not part of the same object (_draw_object, puzzle.py:596-638).
// Every paint stays inside the cell's own pixel square, so a cell's final appearance is fully
// described by the colour of its 3x3 REGIONS (corner / edge / centre; pixels_per_cell >=
// 1 + 2*border_width keeps the regions disjoint and the centre non-empty):
//   region (ry, rx) of a cell occupied by layer L is painted with L's border colour iff
//     (rx == 0 and left absent) or (rx == 2 and right absent) or (ry == 0 and up absent) or
//     (ry == 2 and down absent) or (corner region and the diagonal neighbour absent),
//   else with L's fill colour, else (fill is None: goal outlines) it is left untouched.
// Layers bottom to top: background 255; agent walls (only if the file has 'aw' -- and then the
// layer covers aw UNION w, because the reference's agent-wall object shares its cell set with
// the collision set extended at puzzle.py:273); walls; movables in state order; goal outlines.
//
// cell_code() returns 9 palette nibbles (region ry*3+rx at bits 4*(ry*3+rx)); palette indices
// are listed in pushworld_b200/tables.py.
#pragma once
#include <stdint.h>

#include "pw_core.cuh"

namespace pw {

constexpr uint32_t kPalNone = 15;
constexpr uint32_t kPalAgentWall = 3, kPalAgentWallBorder = 4, kPalWall = 10, kPalWallBorder = 11;
constexpr uint32_t kPalPad = 12;

// 3 occupancy bits (c-1, c, c+1) of a row mask, for 0 <= c <= 63
PW_HD uint32_t bits3(uint64_t row, int c) {
  return (uint32_t)(c == 0 ? (row << 1) : (row >> (c - 1))) & 7u;
}

// Paints one layer over `code`.  nb: 9 neighbourhood bits, bit (dy+1)*3 + (dx+1); the centre
// (bit 4) must be set.
PW_HD uint64_t paint(uint64_t code, uint32_t nb, uint32_t fill, uint32_t border) {
  const uint32_t absent = ~nb;
  const bool l = (absent >> 3) & 1u, r = (absent >> 5) & 1u, u = (absent >> 1) & 1u,
             d = (absent >> 7) & 1u;
  const bool ul = absent & 1u, ur = (absent >> 2) & 1u, dl = (absent >> 6) & 1u,
             dr = (absent >> 8) & 1u;
  const bool is_border[9] = {l || u || ul, u, r || u || ur, l, false, r, l || d || dl, d,
                             r || d || dr};
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const uint32_t col = is_border[k] ? border : fill;
    if (col != kPalNone) code = (code & ~(0xfull << (4 * k))) | ((uint64_t)col << (4 * k));
  }
  return code;
}

// Neighbourhood of grid cell (cx, cy) in a static layer stored as one u64 per grid row.
template <class Table>
PW_HD uint32_t static_nb(const Table& t, int off, int H, int cx, int cy) {
  const uint64_t up = cy > 0 ? t.row(off + cy - 1) : 0ull;
  const uint64_t mid = t.row(off + cy);
  const uint64_t dn = cy + 1 < H ? t.row(off + cy + 1) : 0ull;
  return bits3(up, cx) | (bits3(mid, cx) << 3) | (bits3(dn, cx) << 6);
}

// Neighbourhood of grid cell (cx, cy) in an object with descriptor `d` placed at (px, py);
// returns 0 if the object does not occupy the cell.
template <class Table>
PW_HD uint32_t object_nb(const Table& t, uint32_t d, int px, int py, int cx, int cy) {
  const int lx = cx - px, ly = cy - py;
  const int w = desc_w(d), h = desc_h(d);
  if (lx < 0 || ly < 0 || lx >= w || ly >= h) return 0;
  const int o = desc_off(d);
  const uint64_t mid = t.row(o + ly);
  if (!((mid >> lx) & 1ull)) return 0;
  const uint64_t up = ly > 0 ? t.row(o + ly - 1) : 0ull;
  const uint64_t dn = ly + 1 < h ? t.row(o + ly + 1) : 0ull;
  return bits3(up, lx) | (bits3(mid, lx) << 3) | (bits3(dn, lx) << 6);
}

// Final 9-region colour code of grid cell (cx, cy).  `pos(j)` -> (x, y) of movable j.
template <class Table, class Pos>
PW_HD uint64_t cell_code(const Table& t, const Pos& pos, int cx, int cy) {
  const uint32_t dims = t.word(kHdrDims), counts = t.word(kHdrCounts);
  const int H = (int)(dims >> 16);
  const int n = (int)(counts & 0xffffu), g = (int)(counts >> 16);
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
  uint64_t code = 0;  // palette 0 = background
  uint32_t nb;
  if (t.word(kHdrFlags) & 1u) {
    nb = static_nb(t, (int)t.word(kHdrAwallOff), H, cx, cy);
    if (nb & 16u) code = paint(code, nb, kPalAgentWall, kPalAgentWallBorder);
  }
  nb = static_nb(t, (int)t.word(kHdrWallOff), H, cx, cy);
  if (nb & 16u) code = paint(code, nb, kPalWall, kPalWallBorder);
  for (int j = 0; j < n; ++j) {
    const uint32_t d = t.word(desc_word + 4 * j);
    nb = object_nb(t, d, pos.x(j), pos.y(j), cx, cy);
    if (nb) {
      const uint32_t pal = t.word(desc_word + 4 * j + 3);
      code = paint(code, nb, pal & 0xffu, (pal >> 8) & 0xffu);
    }
  }
  for (int k = 0; k < g; ++k) {
    const int j = n + k;
    const uint32_t d = t.word(desc_word + 4 * j);
    const uint32_t at = t.word(desc_word + 4 * j + 1);
    nb = object_nb(t, d, lo16(at), hi16(at), cx, cy);
    if (nb) {
      const uint32_t pal = t.word(desc_word + 4 * j + 3);
      code = paint(code, nb, pal & 0xffu, (pal >> 8) & 0xffu);
    }
  }
  return code;
}

// region index (0..2) of an in-cell pixel offset
PW_HD int region_of(int o, int ppc, int bw) { return o < bw ? 0 : (o >= ppc - bw ? 2 : 1); }

}  // namespace pw

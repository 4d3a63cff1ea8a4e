// pw_compile.cpp -- native ".pwp text -> table blob" compiler (host only).
//
// Follows PushWorldPuzzle.__init__ of the reference (python3/src/pushworld/puzzle.py:130-257)
// for everything that defines the state layout: whitespace-separated tokens, '+' joins several
// ids in one cell, ids are lower-cased, '.' is empty, file cell (c, r) becomes grid cell
// (c+1, r+1) inside a 1-cell wall border, ids are visited in DESCENDING string order so that
// goal g<k> pulls movable m<k> to the front of the state, remaining m* follow in order of first
// appearance.  Instead of the reference's collision hash sets (puzzle.py:259-308) it emits
// row bitmasks; layout in include/pushworld_b200.h.  The Python twin is
// pushworld_b200/tables.py and tests check both produce identical bytes.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pushworld_b200.h"
#include "pw_host.h"

namespace pw {
namespace {

using Cell = std::pair<int, int>;  // (x, y)
using CellSet = std::set<Cell>;

enum Pal : uint32_t {
  kPalAgent = 1,
  kPalAgentBorder = 2,
  kPalGoalBorder = 5,
  kPalGoalObject = 6,
  kPalGoalObjectBorder = 7,
  kPalMovable = 8,
  kPalMovableBorder = 9,
  kPalNone = 15,
};

bool is_space(char c) {
  return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v' || c == '\n';
}

// Splits like Python's text-mode file iteration: "\n", "\r\n" and a lone "\r" end a line; a
// trailing line terminator does not open another (empty) line.
std::vector<std::string> split_lines(const char* text, size_t len) {
  std::vector<std::string> lines;
  std::string cur;
  bool open = false;
  for (size_t i = 0; i < len; ++i) {
    const char c = text[i];
    if (c == '\n' || c == '\r') {
      if (c == '\r' && i + 1 < len && text[i + 1] == '\n') ++i;
      lines.push_back(cur);
      cur.clear();
      open = false;
    } else {
      cur.push_back(c);
      open = true;
    }
  }
  if (open) lines.push_back(cur);
  return lines;
}

std::vector<std::string> split_tokens(const std::string& line) {  // str.split()
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && is_space(line[i])) ++i;
    size_t j = i;
    while (j < line.size() && !is_space(line[j])) ++j;
    if (j > i) out.push_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

std::vector<uint64_t> rows_of(const CellSet& cells, int n_rows) {
  std::vector<uint64_t> rows((size_t)n_rows, 0);
  for (const Cell& c : cells) rows[(size_t)c.second] |= 1ull << c.first;
  return rows;
}

}  // namespace

int compile_puzzle(const char* text, size_t len, std::vector<uint8_t>* out) {
  // ---- parse (puzzle.py:131-157) ----
  std::map<std::string, CellSet> cells;
  std::vector<std::string> first_seen;  // dict insertion order of the reference
  const std::vector<std::string> lines = split_lines(text, len);
  int n_cols = -1;
  for (size_t r = 0; r < lines.size(); ++r) {
    const std::vector<std::string> tokens = split_tokens(lines[r]);
    if (r == 0) {
      n_cols = (int)tokens.size();
    } else if ((int)tokens.size() != n_cols) {
      return set_error(PW_E_FORMAT,
                       "Row " + std::to_string(r + 1) +
                           " does not have the same number of elements as the first row.");
    }
    for (size_t c = 0; c < tokens.size(); ++c) {
      const std::string& token = tokens[c];
      size_t start = 0;
      while (true) {
        const size_t plus = token.find('+', start);
        std::string name =
            token.substr(start, plus == std::string::npos ? std::string::npos : plus - start);
        for (char& ch : name)
          if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
        if (name.empty()) return set_error(PW_E_FORMAT, "empty object id in a cell");
        if (name != ".") {
          if (cells.find(name) == cells.end()) first_seen.push_back(name);
          cells[name].insert(Cell((int)c + 1, (int)r + 1));
        }
        if (plus == std::string::npos) break;
        start = plus + 1;
      }
    }
  }
  if (cells.find("a") == cells.end())
    return set_error(PW_E_FORMAT, "Every puzzle must have an agent object, indicated by 'a'.");
  if (n_cols <= 0) return set_error(PW_E_FORMAT, "The puzzle has no cells.");

  const int width = n_cols + 2, height = (int)lines.size() + 2;  // puzzle.py:160-161
  if (width > PW_MAX_WIDTH || height > PW_MAX_HEIGHT)
    return set_error(PW_E_LIMIT, "grid " + std::to_string(width) + "x" + std::to_string(height) +
                                     " exceeds the " + std::to_string(PW_MAX_WIDTH) + "x" +
                                     std::to_string(PW_MAX_HEIGHT) + " table limit");
  const bool has_aw = cells.find("aw") != cells.end();
  if (cells.find("w") == cells.end()) first_seen.push_back("w");
  CellSet& walls = cells["w"];
  for (int x = 0; x < width; ++x) {
    walls.insert(Cell(x, 0));
    walls.insert(Cell(x, height - 1));
  }
  for (int y = 0; y < height; ++y) {
    walls.insert(Cell(0, y));
    walls.insert(Cell(width - 1, y));
  }

  // ---- object order (puzzle.py:170-237) ----
  std::vector<std::string> order{"a"}, goals;
  for (auto it = cells.rbegin(); it != cells.rend(); ++it) {  // descending ids
    const std::string& name = it->first;
    if (name[0] == 'g') {
      const std::string movable = "m" + name.substr(1);
      if (cells.find(movable) == cells.end())
        return set_error(PW_E_FORMAT, "Goal has no associated movable object: " + movable);
      goals.push_back(name);
      order.push_back(movable);
    }
  }
  for (const std::string& name : first_seen)
    if (name[0] == 'm' && std::find(order.begin(), order.end(), name) == order.end())
      order.push_back(name);
  const int n = (int)order.size(), g = (int)goals.size();
  if (n > PW_MAX_OBJECTS)
    return set_error(PW_E_LIMIT, std::to_string(n) + " movable objects; the table format allows " +
                                     std::to_string(PW_MAX_OBJECTS));

  struct Shape {
    Cell origin;
    CellSet rel;
    int w, h;
  };
  auto shape_of = [&](const std::string& name) {
    const CellSet& pts = cells[name];
    Shape s;
    int mx = 1 << 30, my = 1 << 30, Mx = -1, My = -1;
    for (const Cell& c : pts) {
      mx = std::min(mx, c.first);
      my = std::min(my, c.second);
      Mx = std::max(Mx, c.first);
      My = std::max(My, c.second);
    }
    s.origin = Cell(mx, my);
    for (const Cell& c : pts) s.rel.insert(Cell(c.first - mx, c.second - my));
    s.w = Mx - mx + 1;
    s.h = My - my + 1;
    return s;
  };

  CellSet agent_walls = walls;  // aw UNION w, puzzle.py:273
  if (has_aw) agent_walls.insert(cells["aw"].begin(), cells["aw"].end());

  // ---- serialise ----
  const int n_desc = n + g;
  const uint32_t obj_off = PW_TABLE_HEADER_WORDS / 2;
  const uint32_t pool_off = obj_off + (uint32_t)n_desc * 2;
  std::vector<uint64_t> pool;
  auto push_rows = [&](const std::vector<uint64_t>& rr) {
    const uint32_t off = pool_off + (uint32_t)pool.size();
    pool.insert(pool.end(), rr.begin(), rr.end());
    return off;
  };
  const uint32_t wall_off = push_rows(rows_of(walls, height));
  const uint32_t awall_off = push_rows(rows_of(agent_walls, height));

  std::vector<uint32_t> desc;
  std::vector<Shape> shapes;
  for (int k = 0; k < n_desc; ++k) {
    const Shape s = shape_of(k < n ? order[(size_t)k] : goals[(size_t)(k - n)]);
    shapes.push_back(s);
    uint32_t goal_word = PW_NO_GOAL, pal;
    if (k < n) {
      if (k >= 1 && k <= g) {
        const Shape gs = shape_of(goals[(size_t)(k - 1)]);
        goal_word = ((uint32_t)gs.origin.first & 0xffffu) | (((uint32_t)gs.origin.second & 0xffffu) << 16);
      }
      pal = k == 0 ? (kPalAgent | (kPalAgentBorder << 8))
                   : (k <= g ? (kPalGoalObject | (kPalGoalObjectBorder << 8))
                             : (kPalMovable | (kPalMovableBorder << 8)));
    } else {
      pal = kPalNone | (kPalGoalBorder << 8);
    }
    const uint32_t off = push_rows(rows_of(s.rel, s.h));
    if (off > 0xffffu) return set_error(PW_E_LIMIT, "puzzle geometry exceeds the 16-bit row offset");
    // the descriptor (and the pair windows of pw_fast.cuh) hold bounding boxes in 8 bits
    if (s.h > 255 || s.w > PW_MAX_WIDTH)
      return set_error(PW_E_LIMIT, "an object's bounding box exceeds " + std::to_string(PW_MAX_WIDTH) +
                                       " columns x 255 rows");
    desc.push_back(off | ((uint32_t)s.w << 16) | ((uint32_t)s.h << 24));
    desc.push_back(((uint32_t)s.origin.first & 0xffffu) | (((uint32_t)s.origin.second & 0xffffu) << 16));
    desc.push_back(goal_word);
    desc.push_back(pal);
  }

  // "regular": no movable overlaps another movable / a wall, agent clear of aw UNION w
  bool regular = true;
  {
    CellSet seen;
    for (int k = 0; k < n && regular; ++k) {
      const CellSet& pts = cells[order[(size_t)k]];
      const CellSet& obstacles = k == 0 ? agent_walls : walls;
      for (const Cell& c : pts)
        if (seen.count(c) || obstacles.count(c)) {
          regular = false;
          break;
        }
      seen.insert(pts.begin(), pts.end());
    }
  }

  const uint32_t size_u64 = pool_off + (uint32_t)pool.size();
  uint32_t header[PW_TABLE_HEADER_WORDS] = {0};
  header[0] = PW_TABLE_MAGIC;
  header[1] = (uint32_t)width | ((uint32_t)height << 16);
  header[2] = (uint32_t)n | ((uint32_t)g << 16);
  header[3] = (has_aw ? PW_TABLE_FLAG_HAS_AW : 0u) | (regular ? PW_TABLE_FLAG_REGULAR : 0u);
  header[4] = wall_off;
  header[5] = awall_off;
  header[6] = obj_off;
  header[7] = size_u64;
  out->assign((size_t)size_u64 * 8, 0);
  uint8_t* p = out->data();
  memcpy(p, header, sizeof(header));
  memcpy(p + sizeof(header), desc.data(), desc.size() * 4);
  memcpy(p + (size_t)pool_off * 8, pool.data(), pool.size() * 8);
  return PW_OK;
}

}  // namespace pw

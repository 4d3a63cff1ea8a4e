// Minimal stand-in for <boost/functional/hash.hpp> -- TEST INFRASTRUCTURE (oracle/_ref build).
// The reference's C++ simulator (cpp/include/pushworld_puzzle.h:22,74-78, cpp/src/pushworld_puzzle.cc:
// 50-56, cpp/include/heuristics/novelty.h:32-38) uses Boost only to hash keys of unordered
// containers; any deterministic hash gives the same program behaviour.  This header provides the
// two entry points it calls, with Boost's published combining formula.
#ifndef PW_REF_SHIM_BOOST_FUNCTIONAL_HASH_HPP_
#define PW_REF_SHIM_BOOST_FUNCTIONAL_HASH_HPP_
// (the real Boost headers pull in these standard headers, which the reference then relies on)
#include <algorithm>
#include <climits>
#include <cstddef>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>

namespace boost {

template <class T>
inline void hash_combine(std::size_t& seed, const T& value) {
  seed ^= std::hash<T>()(value) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}

template <class It>
inline std::size_t hash_range(It first, It last) {
  std::size_t seed = 0;
  for (; first != last; ++first) hash_combine(seed, *first);
  return seed;
}

}  // namespace boost
#endif

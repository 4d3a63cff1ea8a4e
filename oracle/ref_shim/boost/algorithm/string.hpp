// Minimal stand-in for <boost/algorithm/string.hpp> -- TEST INFRASTRUCTURE (oracle/_ref build).
// Provides exactly what cpp/src/pushworld_puzzle.cc:207-226 calls: trim, split with
// is_any_of + token_compress_on, to_lower -- with Boost's documented behaviour (split keeps empty
// leading / trailing tokens; token_compress_on merges adjacent separators; trim strips
// std::isspace characters from both ends).
#ifndef PW_REF_SHIM_BOOST_ALGORITHM_STRING_HPP_
#define PW_REF_SHIM_BOOST_ALGORITHM_STRING_HPP_
#include <cctype>
#include <string>

namespace boost {

enum token_compress_mode_type { token_compress_on, token_compress_off };

struct is_any_of {
  std::string chars;
  explicit is_any_of(const char* c) : chars(c) {}
  bool operator()(char ch) const { return chars.find(ch) != std::string::npos; }
};

inline void trim(std::string& s) {
  std::size_t b = 0, e = s.size();
  while (b < e && std::isspace(static_cast<unsigned char>(s[b]))) ++b;
  while (e > b && std::isspace(static_cast<unsigned char>(s[e - 1]))) --e;
  s = s.substr(b, e - b);
}

inline void to_lower(std::string& s) {
  for (char& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
}

template <class Container, class Pred>
inline Container& split(Container& result, const std::string& input, Pred is_separator,
                        token_compress_mode_type compress = token_compress_off) {
  result.clear();
  std::string token;
  std::size_t i = 0;
  const std::size_t n = input.size();
  while (true) {
    token.clear();
    while (i < n && !is_separator(input[i])) token.push_back(input[i++]);
    result.push_back(token);
    if (i >= n) break;
    ++i;  // the separator
    if (compress == token_compress_on)
      while (i < n && is_separator(input[i])) ++i;
  }
  return result;
}

}  // namespace boost
#endif

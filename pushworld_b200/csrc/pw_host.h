// pw_host.h -- host-side helpers shared by the C-ABI translation units (not public).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pw {

// Records a thread-local error message and returns `code` (so callers can `return set_error(..)`).
int set_error(int code, const std::string& message);
void clear_error();

// .pwp text -> table blob (pw_compile.cpp)
int compile_puzzle(const char* text, size_t len, std::vector<uint8_t>* out);

}  // namespace pw

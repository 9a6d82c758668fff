// Boost shim (oracle build only): function.hpp
// The reference uses nothing from Boost.Function, but relies on standard headers that real
// Boost headers pull in transitively (assert, chrono, unordered_map ...).
#ifndef SHIM_FUNCTION_HPP
#define SHIM_FUNCTION_HPP
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <ctime>
#include <deque>
#include <functional>
#include <unordered_map>
#include <unordered_set>
#endif

// Boost shim (oracle build only): algorithm/string/predicate.hpp (ends_with),
// used at /root/reference/config.h:260 and build.h:122.
#ifndef SHIM_BOOST_ALGO_PREDICATE_HPP
#define SHIM_BOOST_ALGO_PREDICATE_HPP
#include <string>
namespace boost {
namespace algorithm {
inline bool ends_with(const std::string& s, const std::string& suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
}  // namespace algorithm
using algorithm::ends_with;
}  // namespace boost
#endif

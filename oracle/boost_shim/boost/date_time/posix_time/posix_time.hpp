// Boost shim (oracle build only): date_time/posix_time (second_clock, to_iso_string); only
// referenced by a function the reference never calls (/root/reference/fora.cpp:39-47).
#ifndef SHIM_BOOST_POSIX_TIME_HPP
#define SHIM_BOOST_POSIX_TIME_HPP
#include <ctime>
#include <string>
namespace boost {
namespace posix_time {
struct ptime {
    std::time_t t;
};
struct second_clock {
    static ptime local_time() { return ptime{std::time(nullptr)}; }
};
inline std::string to_iso_string(const ptime& p) {
    char buf[32];
    std::strftime(buf, sizeof buf, "%Y%m%dT%H%M%S", std::localtime(&p.t));
    return buf;
}
}  // namespace posix_time
}  // namespace boost
#endif

// Boost shim (oracle build only): filesystem.hpp (exists / path / create_directories),
// used at /root/reference/config.h:263-266.
#ifndef SHIM_BOOST_FILESYSTEM_HPP
#define SHIM_BOOST_FILESYSTEM_HPP
#include <string>
#include <sys/stat.h>
#include <sys/types.h>

namespace boost {
namespace filesystem {
class path {
    std::string s_;

public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    const std::string& string() const { return s_; }
};
inline bool exists(const path& p) {
    struct stat st;
    return ::stat(p.string().c_str(), &st) == 0;
}
inline bool create_directories(const path& p) {
    const std::string& s = p.string();
    bool made = false;
    for (size_t i = 1; i <= s.size(); ++i) {
        if (i == s.size() || s[i] == '/') {
            std::string sub = s.substr(0, i);
            if (!sub.empty() && !exists(path(sub))) made = (::mkdir(sub.c_str(), 0777) == 0) || made;
        }
    }
    return made;
}
}  // namespace filesystem
}  // namespace boost
#endif

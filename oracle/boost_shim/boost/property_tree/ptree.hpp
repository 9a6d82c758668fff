// Boost shim (oracle build only): property_tree/ptree.hpp + write_json.
// TEST INFRASTRUCTURE. Ordered string tree with put / put_child and a JSON writer that quotes
// every leaf, as Boost.PropertyTree does (/root/reference/config.h:140-159,198-230,281-307).
#ifndef SHIM_BOOST_PTREE_HPP
#define SHIM_BOOST_PTREE_HPP
#include <limits>
#include <ostream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace boost {
namespace property_tree {

class ptree {
public:
    std::string data;
    std::vector<std::pair<std::string, ptree> > children;

    ptree* find(const std::string& k) {
        for (auto& c : children)
            if (c.first == k) return &c.second;
        return nullptr;
    }
    ptree& walk(const std::string& path) {
        ptree* cur = this;
        size_t pos = 0;
        while (true) {
            size_t dot = path.find('.', pos);
            std::string k = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            ptree* nxt = cur->find(k);
            if (!nxt) {
                cur->children.push_back(std::make_pair(k, ptree()));
                nxt = &cur->children.back().second;
            }
            cur = nxt;
            if (dot == std::string::npos) break;
            pos = dot + 1;
        }
        return *cur;
    }
    template <class T>
    ptree& put(const std::string& path, const T& v) {
        std::ostringstream ss;
        ss.precision(std::numeric_limits<double>::max_digits10);
        ss << std::boolalpha << v;
        ptree& t = walk(path);
        t.data = ss.str();
        return t;
    }
    ptree& put_child(const std::string& path, const ptree& child) {
        ptree& t = walk(path);
        t = child;
        return t;
    }
};

namespace shim_detail {
inline std::string esc(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '/') o += "\\/";
        else if (c == '\n') o += "\\n";
        else if (c == '\t') o += "\\t";
        else o += c;
    }
    return o;
}
inline void write(std::ostream& os, const ptree& t, int indent, bool pretty) {
    if (t.children.empty()) {
        os << '"' << esc(t.data) << '"';
        return;
    }
    os << '{';
    if (pretty) os << '\n';
    for (size_t i = 0; i < t.children.size(); ++i) {
        if (pretty) os << std::string(4 * (indent + 1), ' ');
        os << '"' << esc(t.children[i].first) << "\":";
        if (pretty) os << ' ';
        write(os, t.children[i].second, indent + 1, pretty);
        if (i + 1 < t.children.size()) os << ',';
        if (pretty) os << '\n';
    }
    if (pretty) os << std::string(4 * indent, ' ');
    os << '}';
}
}  // namespace shim_detail

inline void write_json(std::ostream& os, const ptree& t, bool pretty = true) {
    shim_detail::write(os, t, 0, pretty);
    os << '\n';
}

}  // namespace property_tree
}  // namespace boost
#endif

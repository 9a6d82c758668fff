// Boost shim (oracle build only): archive classes.
// TEST INFRASTRUCTURE. Stand-in for boost::archive::{binary,text}_{o,i}archive as used at
// /root/reference/build.h:53-80,130-131,141-142,199-206,211-216. Emits the layout described
// in SURVEY.md Appendix A (x86-64 Boost.Serialization, library version 15): that layout is
// written from knowledge of the format, NOT verified against a real Boost build.
#ifndef SHIM_BOOST_ARCHIVE_HPP
#define SHIM_BOOST_ARCHIVE_HPP
#include <cstdint>
#include <cstring>
#include <iomanip>
#include <istream>
#include <limits>
#include <map>
#include <ostream>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <vector>

namespace boost {
namespace archive {

namespace shim_detail {
template <class T>
struct is_bitwise : std::is_arithmetic<T> {};
template <class A, class B>
struct is_bitwise<std::pair<A, B> >
    : std::integral_constant<bool, is_bitwise<A>::value && is_bitwise<B>::value> {};
}  // namespace shim_detail

// ------------------------------------------------------------------ binary
class binary_oarchive {
    std::ostream& os_;
    std::set<std::string> seen_;
    void raw(const void* p, size_t n) { os_.write(static_cast<const char*>(p), std::streamsize(n)); }
    template <class T>
    void pod(const T& v) { raw(&v, sizeof(T)); }
    template <class T>
    void class_info() {
        if (seen_.insert(typeid(T).name()).second) {
            pod<uint8_t>(0);   // tracking
            pod<uint32_t>(0);  // class version
        }
    }
    template <class T>
    typename std::enable_if<std::is_arithmetic<T>::value>::type save(const T& v) { pod(v); }
    template <class A, class B>
    void save(const std::pair<A, B>& p) {
        class_info<std::pair<A, B> >();
        save(p.first);
        save(p.second);
    }
    template <class T>
    void save(const std::vector<T>& v) {
        class_info<std::vector<T> >();
        pod<uint64_t>(v.size());
        if (shim_detail::is_bitwise<T>::value) {
            if (!v.empty()) raw(v.data(), v.size() * sizeof(T));
        } else {
            pod<uint32_t>(0);  // item version
            for (const T& x : v) save(x);
        }
    }
    template <class K, class V>
    void save(const std::map<K, V>& m) {
        class_info<std::map<K, V> >();
        pod<uint64_t>(m.size());
        pod<uint32_t>(0);  // item version
        for (const auto& kv : m) save(std::pair<K, V>(kv.first, kv.second));
    }

public:
    explicit binary_oarchive(std::ostream& os) : os_(os) {
        const char sig[] = "serialization::archive";
        pod<uint64_t>(22);
        raw(sig, 22);
        pod<uint16_t>(15);
        const uint8_t sizes[4] = {sizeof(int), sizeof(long), sizeof(float), sizeof(double)};
        raw(sizes, 4);
        pod<uint32_t>(1);
    }
    template <class T>
    binary_oarchive& operator<<(const T& v) {
        save(v);
        return *this;
    }
};

class binary_iarchive {
    std::istream& is_;
    std::set<std::string> seen_;
    void raw(void* p, size_t n) {
        is_.read(static_cast<char*>(p), std::streamsize(n));
        if (size_t(is_.gcount()) != n) throw std::runtime_error("shim binary_iarchive: short read");
    }
    template <class T>
    T pod() {
        T v;
        raw(&v, sizeof(T));
        return v;
    }
    template <class T>
    void class_info() {
        if (seen_.insert(typeid(T).name()).second) {
            pod<uint8_t>();
            pod<uint32_t>();
        }
    }
    template <class T>
    typename std::enable_if<std::is_arithmetic<T>::value>::type load(T& v) { v = pod<T>(); }
    template <class A, class B>
    void load(std::pair<A, B>& p) {
        class_info<std::pair<A, B> >();
        load(p.first);
        load(p.second);
    }
    template <class T>
    void load(std::vector<T>& v) {
        class_info<std::vector<T> >();
        uint64_t n = pod<uint64_t>();
        v.resize(n);
        if (shim_detail::is_bitwise<T>::value) {
            if (n) raw(v.data(), n * sizeof(T));
        } else {
            pod<uint32_t>();
            for (T& x : v) load(x);
        }
    }
    template <class K, class V>
    void load(std::map<K, V>& m) {
        class_info<std::map<K, V> >();
        uint64_t n = pod<uint64_t>();
        pod<uint32_t>();
        m.clear();
        for (uint64_t i = 0; i < n; ++i) {
            std::pair<K, V> kv;
            load(kv);
            m.insert(kv);
        }
    }

public:
    explicit binary_iarchive(std::istream& is) : is_(is) {
        uint64_t len = pod<uint64_t>();
        if (len != 22) throw std::runtime_error("shim binary_iarchive: bad signature length");
        char sig[22];
        raw(sig, 22);
        pod<uint16_t>();
        uint8_t sizes[4];
        raw(sizes, 4);
        pod<uint32_t>();
    }
    template <class T>
    binary_iarchive& operator>>(T& v) {
        load(v);
        return *this;
    }
};

// ------------------------------------------------------------------ text
class text_oarchive {
    std::ostream& os_;
    std::set<std::string> seen_;
    template <class T>
    void class_info() {
        if (seen_.insert(typeid(T).name()).second) os_ << " 0 0";
    }
    template <class T>
    typename std::enable_if<std::is_integral<T>::value>::type save(const T& v) { os_ << ' ' << v; }
    void save(const double& v) {
        os_ << ' ' << std::setprecision(17) << std::scientific << v;
        os_.unsetf(std::ios_base::floatfield);
    }
    template <class A, class B>
    void save(const std::pair<A, B>& p) {
        class_info<std::pair<A, B> >();
        save(p.first);
        save(p.second);
    }
    template <class T>
    void save(const std::vector<T>& v) {
        class_info<std::vector<T> >();
        os_ << ' ' << v.size() << " 0";
        for (const T& x : v) save(x);
    }
    template <class K, class V>
    void save(const std::map<K, V>& m) {
        class_info<std::map<K, V> >();
        os_ << ' ' << m.size() << " 0";
        for (const auto& kv : m) save(std::pair<K, V>(kv.first, kv.second));
    }

public:
    explicit text_oarchive(std::ostream& os) : os_(os) { os_ << "22 serialization::archive 15"; }
    ~text_oarchive() { os_ << '\n'; }
    template <class T>
    text_oarchive& operator<<(const T& v) {
        save(v);
        return *this;
    }
};

class text_iarchive {
    std::istream& is_;
    std::set<std::string> seen_;
    template <class T>
    void class_info() {
        if (seen_.insert(typeid(T).name()).second) {
            int a, b;
            is_ >> a >> b;
        }
    }
    template <class T>
    typename std::enable_if<std::is_arithmetic<T>::value>::type load(T& v) { is_ >> v; }
    template <class A, class B>
    void load(std::pair<A, B>& p) {
        class_info<std::pair<A, B> >();
        load(p.first);
        load(p.second);
    }
    template <class T>
    void load(std::vector<T>& v) {
        class_info<std::vector<T> >();
        size_t n;
        int item_version;
        is_ >> n >> item_version;
        v.resize(n);
        for (T& x : v) load(x);
    }
    template <class K, class V>
    void load(std::map<K, V>& m) {
        class_info<std::map<K, V> >();
        size_t n;
        int item_version;
        is_ >> n >> item_version;
        m.clear();
        for (size_t i = 0; i < n; ++i) {
            std::pair<K, V> kv;
            load(kv);
            m.insert(kv);
        }
    }

public:
    explicit text_iarchive(std::istream& is) : is_(is) {
        int len, ver;
        std::string sig;
        is_ >> len >> sig >> ver;
        if (sig != "serialization::archive") throw std::runtime_error("shim text_iarchive: bad signature");
    }
    template <class T>
    text_iarchive& operator>>(T& v) {
        load(v);
        return *this;
    }
};

}  // namespace archive
}  // namespace boost
#endif

// fora_b200/host/archive.hpp -- reader/writer for the three files the reference persists through
// Boost.Serialization (/root/reference/build.h:121-145,194-217):
//   randwalks.idx[.onehopopt]    binary_oarchive of std::vector<int>
//   randwalks.info[.onehopopt]   binary_oarchive of std::vector<std::pair<unsigned long long, unsigned long>>
//   <dataset>.topk.pprs          text_oarchive   of std::map<int, std::vector<std::pair<int,double>>>
// Layout as described in SURVEY.md Appendix A (x86-64, library version 15).  Boost is not available in
// this environment, so byte-for-byte interop with a real Boost build is unverified; the enforced
// contract is round-trip self-consistency (only ./fora itself reads these files).  The reader accepts
// any library version >= 6.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fora_host {

struct BinHeader {
    static void write(std::ostream& os) {
        const uint64_t len = 22;
        os.write((const char*)&len, 8);
        os.write("serialization::archive", 22);
        const uint16_t ver = 15;
        os.write((const char*)&ver, 2);
        const uint8_t sizes[4] = {sizeof(int), sizeof(long), sizeof(float), sizeof(double)};
        os.write((const char*)sizes, 4);
        const uint32_t endian = 1;
        os.write((const char*)&endian, 4);
        const uint8_t tracking = 0;      // class info of the top-level object
        const uint32_t class_version = 0;
        os.write((const char*)&tracking, 1);
        os.write((const char*)&class_version, 4);
    }
    static void read(std::istream& is, const std::string& what) {
        uint64_t len = 0;
        is.read((char*)&len, 8);
        char sig[22];
        if (!is || len != 22) throw std::runtime_error(what + ": not a Boost binary archive");
        is.read(sig, 22);
        if (std::memcmp(sig, "serialization::archive", 22) != 0) throw std::runtime_error(what + ": bad archive signature");
        uint16_t ver = 0;
        is.read((char*)&ver, 2);
        if (ver < 6) throw std::runtime_error(what + ": archive library version too old");
        uint8_t sizes[4];
        is.read((char*)sizes, 4);
        if (sizes[0] != sizeof(int) || sizes[1] != sizeof(long) || sizes[3] != sizeof(double))
            throw std::runtime_error(what + ": archive written on an incompatible platform");
        uint32_t endian = 0;
        is.read((char*)&endian, 4);
        uint8_t tracking;
        uint32_t class_version;
        is.read((char*)&tracking, 1);
        is.read((char*)&class_version, 4);
        if (!is) throw std::runtime_error(what + ": truncated archive header");
    }
};

// rw_idx: flat destinations
inline void save_index_dest(const std::string& path, const std::vector<int32_t>& dest) {
    std::ofstream os(path, std::ios::binary);
    if (!os) throw std::runtime_error("cannot write " + path);
    BinHeader::write(os);
    const uint64_t count = dest.size();
    os.write((const char*)&count, 8);
    if (count) os.write((const char*)dest.data(), (std::streamsize)(count * 4));
}
inline void load_index_dest(const std::string& path, std::vector<int32_t>& dest) {
    std::ifstream is(path, std::ios::binary);
    if (!is) throw std::runtime_error("index file " + path + " not find ");
    BinHeader::read(is, path);
    uint64_t count = 0;
    is.read((char*)&count, 8);
    dest.resize(count);
    if (count) is.read((char*)dest.data(), (std::streamsize)(count * 4));
    if (!is) throw std::runtime_error(path + ": truncated");
}
// rw_idx_info: (offset, count) per node, 16 bytes each
inline void save_index_info(const std::string& path, const std::vector<uint64_t>& off, const std::vector<uint64_t>& cnt) {
    std::ofstream os(path, std::ios::binary);
    if (!os) throw std::runtime_error("cannot write " + path);
    BinHeader::write(os);
    const uint64_t count = off.size();
    os.write((const char*)&count, 8);
    for (uint64_t i = 0; i < count; ++i) {
        os.write((const char*)&off[i], 8);
        os.write((const char*)&cnt[i], 8);
    }
}
inline void load_index_info(const std::string& path, std::vector<uint64_t>& off, std::vector<uint64_t>& cnt) {
    std::ifstream is(path, std::ios::binary);
    if (!is) throw std::runtime_error("index file " + path + " not find ");
    BinHeader::read(is, path);
    uint64_t count = 0;
    is.read((char*)&count, 8);
    off.resize(count);
    cnt.resize(count);
    for (uint64_t i = 0; i < count; ++i) {
        is.read((char*)&off[i], 8);
        is.read((char*)&cnt[i], 8);
    }
    if (!is) throw std::runtime_error(path + ": truncated");
}

// exact top-k: text archive of map<int, vector<pair<int,double>>>
typedef std::map<int, std::vector<std::pair<int, double> > > ExactTopk;
inline void save_exact_topk(const std::string& path, const ExactTopk& m) {
    std::ofstream os(path);
    if (!os) throw std::runtime_error("cannot write " + path);
    os << "22 serialization::archive 15 0 0 " << m.size() << " 0";
    bool first_pair = true, first_vec = true, first_inner = true;
    for (const auto& kv : m) {
        if (first_pair) { os << " 0 0"; first_pair = false; }
        os << ' ' << kv.first;
        if (first_vec) { os << " 0 0"; first_vec = false; }
        os << ' ' << kv.second.size() << " 0";
        for (const auto& p : kv.second) {
            if (first_inner) { os << " 0 0"; first_inner = false; }
            os << ' ' << p.first << ' ' << std::setprecision(17) << std::scientific << p.second;
            os.unsetf(std::ios_base::floatfield);
        }
    }
    os << '\n';
}
inline bool load_exact_topk(const std::string& path, ExactTopk& m) {
    std::ifstream is(path);
    if (!is) return false;
    int len, ver, a, b;
    std::string sig;
    is >> len >> sig >> ver;
    if (sig != "serialization::archive") throw std::runtime_error(path + ": not a Boost text archive");
    size_t count;
    int item_version;
    is >> a >> b >> count >> item_version;
    bool first_pair = true, first_vec = true, first_inner = true;
    m.clear();
    for (size_t i = 0; i < count; ++i) {
        if (first_pair) { is >> a >> b; first_pair = false; }
        int key;
        is >> key;
        if (first_vec) { is >> a >> b; first_vec = false; }
        size_t vc;
        is >> vc >> item_version;
        std::vector<std::pair<int, double> > v(vc);
        for (size_t j = 0; j < vc; ++j) {
            if (first_inner) { is >> a >> b; first_inner = false; }
            is >> v[j].first >> v[j].second;
        }
        m[key] = v;
    }
    if (!is) throw std::runtime_error(path + ": truncated text archive");
    return true;
}

} // namespace fora_host

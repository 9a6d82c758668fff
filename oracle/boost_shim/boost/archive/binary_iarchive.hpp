// Boost shim (oracle build only): archive/binary_iarchive.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_ARCHIVE_BINARY_IARCHIVE_HPP
#define SHIM_ARCHIVE_BINARY_IARCHIVE_HPP
#include <boost/archive/text_oarchive.hpp>
#endif

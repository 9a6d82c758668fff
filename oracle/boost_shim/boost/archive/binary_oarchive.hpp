// Boost shim (oracle build only): archive/binary_oarchive.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_ARCHIVE_BINARY_OARCHIVE_HPP
#define SHIM_ARCHIVE_BINARY_OARCHIVE_HPP
#include <boost/archive/text_oarchive.hpp>
#endif

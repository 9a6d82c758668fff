// Boost shim (oracle build only): serialization/utility.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_SERIALIZATION_UTILITY_HPP
#define SHIM_SERIALIZATION_UTILITY_HPP
#include <boost/archive/text_oarchive.hpp>
#endif

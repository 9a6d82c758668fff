// Boost shim (oracle build only): serialization/serialization.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_SERIALIZATION_SERIALIZATION_HPP
#define SHIM_SERIALIZATION_SERIALIZATION_HPP
#include <boost/archive/text_oarchive.hpp>
#endif

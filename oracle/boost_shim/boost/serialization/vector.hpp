// Boost shim (oracle build only): serialization/vector.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_SERIALIZATION_VECTOR_HPP
#define SHIM_SERIALIZATION_VECTOR_HPP
#include <boost/archive/text_oarchive.hpp>
#endif

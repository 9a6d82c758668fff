// Boost shim (oracle build only): bind.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_BIND_HPP
#define SHIM_BIND_HPP
#endif

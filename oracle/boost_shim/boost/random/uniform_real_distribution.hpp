// Boost shim (oracle build only): random/uniform_real_distribution.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_RANDOM_UNIFORM_REAL_DISTRIBUTION_HPP
#define SHIM_RANDOM_UNIFORM_REAL_DISTRIBUTION_HPP
#endif

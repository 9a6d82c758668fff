// Boost shim (oracle build only): random/normal_distribution.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_RANDOM_NORMAL_DISTRIBUTION_HPP
#define SHIM_RANDOM_NORMAL_DISTRIBUTION_HPP
#endif

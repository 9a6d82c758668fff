// Boost shim (oracle build only): random/mersenne_twister.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_RANDOM_MERSENNE_TWISTER_HPP
#define SHIM_RANDOM_MERSENNE_TWISTER_HPP
#endif

// Boost shim (oracle build only): progress.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_PROGRESS_HPP
#define SHIM_PROGRESS_HPP
#endif

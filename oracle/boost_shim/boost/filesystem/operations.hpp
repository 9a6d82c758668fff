// Boost shim (oracle build only): filesystem/operations.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_FILESYSTEM_OPERATIONS_HPP
#define SHIM_FILESYSTEM_OPERATIONS_HPP
#include <boost/filesystem.hpp>
#endif

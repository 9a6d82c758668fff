// Boost shim (oracle build only): filesystem/path.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_FILESYSTEM_PATH_HPP
#define SHIM_FILESYSTEM_PATH_HPP
#include <boost/filesystem.hpp>
#endif

// Boost shim (oracle build only): program_options.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_PROGRAM_OPTIONS_HPP
#define SHIM_PROGRAM_OPTIONS_HPP
namespace boost { namespace program_options {} }
#endif

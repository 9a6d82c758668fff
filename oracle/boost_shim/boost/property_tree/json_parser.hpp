// Boost shim (oracle build only): property_tree/json_parser.hpp
// Minimal stand-in so the unmodified reference compiles without Boost; see oracle/README.md.
#ifndef SHIM_PROPERTY_TREE_JSON_PARSER_HPP
#define SHIM_PROPERTY_TREE_JSON_PARSER_HPP
#include <boost/property_tree/ptree.hpp>
#endif

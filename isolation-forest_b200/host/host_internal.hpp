// Internal declarations of libifb200_host.so.
#pragma once

#include <string>
#include <vector>

#include "ifb200_host.hpp"
#include "json_min.hpp"

namespace ifb200 {
namespace avro {
ForestTables read_tables(const std::string &dir, bool extended);
void write_tables(const std::string &dir, const ForestTables &t, const std::string &codec);
}  // namespace avro
}  // namespace ifb200

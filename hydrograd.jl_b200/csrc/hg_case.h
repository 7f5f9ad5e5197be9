// The named-array container behind hg_case_* (include/hydrograd_b200.h): what the SRH-2D reader (hg_srh.cpp) and the
// partitioner (hg_partition.cpp) hand back to the caller.  Internal: not part of the ABI.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct hg_case {
  std::map<std::string, std::vector<double>> f64;
  std::map<std::string, std::vector<int64_t>> i64;
  std::map<std::string, std::vector<uint8_t>> u8;
  int64_t dims[16] = {0};
  std::string err;
};

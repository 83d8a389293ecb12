// Opaque handle types shared by the host-side translation units of the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct dnab_decoded_set {
  std::vector<std::string> names, seqs;
  std::vector<double> loglike;
  std::vector<int32_t> status;
};

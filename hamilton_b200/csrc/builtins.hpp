// builtins.hpp — tapes of the built-in fixture systems (see builtins.cpp).
#pragma once
#include <vector>

#include "sysgen.hpp"

namespace hb {
const char* builtin_name(int id);
bool builtin_spec(int id, SystemSpec& spec);
bool builtin_params(int id, const double* user, int n_user, std::vector<double>& tape_params);
}  // namespace hb

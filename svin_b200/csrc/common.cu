#include "common.hpp"

namespace svin {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }
}  // namespace svin

extern "C" {
const char* svin_last_error(void) { return svin::last_error().c_str(); }
const char* svin_version(void) { return "svin_b200 0.1 (sm_100a, fp64 BA + BRISK-2 front-end)"; }
}

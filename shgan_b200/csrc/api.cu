// Library-level entry points: ABI version, thread-local error string, launch counter.
#include "common.cuh"

namespace shgan {
static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launch_count{0};
void set_error(const std::string& msg) { t_last_error = msg; }
}  // namespace shgan

extern "C" int shgan_abi_version(void) { return SHGAN_B200_ABI_VERSION; }
extern "C" const char* shgan_last_error(void) { return shgan::t_last_error.c_str(); }
extern "C" uint64_t shgan_launch_count(void) { return shgan::g_launch_count.load(); }

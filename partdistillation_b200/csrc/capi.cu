// Library-wide state of libpdb200.so: last-error string, launch counter, ABI version.
#include "common.cuh"

namespace pdb {
thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace pdb

extern "C" int pdb_abi_version(void) { return 1; }
extern "C" const char* pdb_last_error(void) { return pdb::g_last_error; }
extern "C" int64_t pdb_launch_count(void) { return pdb::g_launches.load(std::memory_order_relaxed); }

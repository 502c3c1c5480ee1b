#include "common.h"
#include <atomic>
namespace vlo {
static thread_local std::string g_err;
void set_error(const std::string& m){ g_err = m; }
const char* last_error(){ return g_err.c_str(); }
int fail(const std::string& m){ set_error(m); return -1; }
static std::atomic<long long> g_launches{0};
void count_launch(int n){ g_launches += n; }
long long launch_count(){ return g_launches.load(); }
}

// lbc_fast.cu -- sm_100a fast kernels behind the hooks of lbc_fast.h.
#include "lbc_fast.h"

namespace lbc {
long long g_launches = 0;
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;
namespace fast {

static bool g_enabled = true;
bool enabled() { return g_enabled; }
void set_enabled(bool on) { g_enabled = on; }


}  // namespace fast
}  // namespace lbc

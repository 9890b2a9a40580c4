// RESIDENT step engine (placeholder until the cluster kernel lands): reports "not applicable".
#include "jj_host.h"
namespace jj {
int resident_supported(JJHandle*, std::string& why) { why = "not built yet"; return 0; }
int resident_prepare(JJHandle* h) { h->err = "resident engine not built"; return JJ_EINVAL; }
int resident_run(JJHandle* h, long long, int, const long long*, const long long*) { h->err = "resident engine not built"; return JJ_EINVAL; }
int resident_set_state(JJHandle*, const double*, const double*) { return JJ_OK; }
int resident_get_state(JJHandle* h, double*, double*) { h->err = "resident engine not built"; return JJ_EINVAL; }
void resident_free(JJHandle*) {}
}

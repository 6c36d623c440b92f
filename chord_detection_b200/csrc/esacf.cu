#include "common.cuh"
struct EsacfPlan {};
void cdb_free_esacf_plans(cdb_handle* h) { for (auto& kv : h->esacf_plans) delete kv.second; h->esacf_plans.clear(); }
extern "C" int cdb_esacf_chroma(cdb_handle* h, const cdb_esacf_params*, const float*, int64_t, int64_t, int64_t, double*, double*, double*, double*, int, void*) {
  return cdb_fail(h, CDB_E_UNSUPPORTED, "esacf: not built yet");
}

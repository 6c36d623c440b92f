#include "common.cuh"
struct IterF0Plan {};
void cdb_free_iterf0_plans(cdb_handle* h) { for (auto& kv : h->iterf0_plans) delete kv.second; h->iterf0_plans.clear(); }
extern "C" int64_t cdb_iterf0_workspace_bytes(const cdb_iterf0_params*, int64_t, int64_t) { return 0; }
extern "C" int cdb_iterf0_chroma(cdb_handle* h, const cdb_iterf0_params*, const float*, int64_t, int64_t, int64_t, void*, int64_t, double*, double*, double*, double*, int, void*) {
  return cdb_fail(h, CDB_E_UNSUPPORTED, "iterf0: not built yet");
}

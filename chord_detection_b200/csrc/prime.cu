#include "common.cuh"
struct PrimePlan {};
void cdb_free_prime_plans(cdb_handle* h) { for (auto& kv : h->prime_plans) delete kv.second; h->prime_plans.clear(); }
extern "C" int cdb_prime_window_sizes(const cdb_prime_params*, int*) { return CDB_E_UNSUPPORTED; }
extern "C" int cdb_prime_chroma(cdb_handle* h, const cdb_prime_params*, const float*, int64_t, int64_t, int64_t, double*, double*, double*, int, void*) {
  return cdb_fail(h, CDB_E_UNSUPPORTED, "prime: not built yet");
}
extern "C" int cdb_pack_and_key(cdb_handle* h, const double*, int64_t, uint8_t*, int32_t*, void*) {
  return cdb_fail(h, CDB_E_UNSUPPORTED, "pack_and_key: not built yet");
}

// TEST-ONLY: stand-ins for the CUDA-side symbols of libgrimb200.so, so that the host text pipeline
// (csrc/grimb_text.cpp) can be built alone with -fsanitize=address,undefined (run_sanitizers.sh).
#include <cstddef>
#include <cstdint>
struct GrimbEngine;
struct GrimbConfig;
struct GrimbBatch;
struct GrimbResults;
extern "C" void* grimb_pinned_alloc(size_t) { return nullptr; }
extern "C" void grimb_pinned_free(void*) {}
extern "C" int grimb_impute_host(GrimbEngine*, const GrimbConfig*, const GrimbBatch*, GrimbResults*) { return -1; }
extern "C" int grimb_abi_version(void) { return 5; }
static thread_local char g_msg[256] = "sanitizer build: no device code";
extern "C" const char* grimb_last_error(void) { return g_msg; }
extern "C" void grimb_set_error(const char* m) {
  size_t i = 0;
  for (; m && m[i] && i + 1 < sizeof(g_msg); ++i) g_msg[i] = m[i];
  g_msg[i] = 0;
}
#define STUB(name) extern "C" int name(void) { return -1; }
STUB(grimb_tables_build) STUB(grimb_tables_free) STUB(grimb_tables_info) STUB(grimb_tables_export)
STUB(grimb_tables_image_size) STUB(grimb_tables_image_ptr) STUB(grimb_tables_image_copy) STUB(grimb_tables_from_image)
STUB(grimb_engine_create) STUB(grimb_engine_free) STUB(grimb_engine_launches) STUB(grimb_engine_kernel_ms)
STUB(grimb_impute_device) STUB(grimb_impute_device_async) STUB(grimb_impute_finish)
STUB(grimb_tables_build_launches) STUB(grimb_tables_build_ms)

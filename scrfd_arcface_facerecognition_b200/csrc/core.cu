// Library-wide state of libb2f.so: last-error text, launch counter, tuning knobs.
#include "b2f_common.cuh"
#include "../../include/b2f.h"

#include <atomic>
#include <stdarg.h>

namespace b2f {
std::atomic<long long> g_launches{0};
extern int g_smem_budget_single;
extern int g_max_block_n;
extern int g_persistent;
extern int g_vhalo;
extern int g_debug;
extern int g_a_res;
extern int g_match_pair;
extern int g_tile_groups, g_tile_mt, g_tile_amode, g_tile_epi, g_tile_max_n, g_tile_cg2, g_tile_cg2_min_n, g_tile_pdl, g_tile_wide_res, g_tile_big_res, g_tile_reduce;
}  // namespace b2f

static thread_local char g_err[1024] = "";

void b2f_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int b2f_version(void) { return B2F_ABI_VERSION; }
extern "C" const char* b2f_last_error(void) { return g_err; }
extern "C" long long b2f_launch_count(void) { return b2f::g_launches.load(); }

extern "C" int b2f_set_tuning(int key, int value) {
  if (key == 0) {
    B2F_REQUIRE(value >= 48 * 1024 && value <= 224 * 1024, "tuning 0 (smem budget) out of range: %d", value);
    b2f::g_smem_budget_single = value;
    return 0;
  }
  if (key == 1) {
    B2F_REQUIRE(value >= 16 && value <= 256 && value % 16 == 0, "tuning 1 (max UMMA N) out of range: %d", value);
    b2f::g_max_block_n = value;
    return 0;
  }
  if (key == 2) {
    b2f::g_persistent = value;
    return 0;
  }
  if (key == 3) {
    b2f::g_vhalo = value;
    return 0;
  }
  if (key == 4) {
    b2f::g_debug = value;
    return 0;
  }
  if (key == 5) { b2f::g_tile_groups = value; return 0; }
  if (key == 6) { b2f::g_tile_mt = value; return 0; }
  if (key == 7) { b2f::g_tile_amode = value; return 0; }
  if (key == 8) { b2f::g_tile_epi = value; return 0; }
  if (key == 9) { b2f::g_tile_max_n = value; return 0; }
  if (key == 10) { b2f::g_a_res = value; return 0; }
  if (key == 11) { b2f::g_tile_cg2 = value; return 0; }
  if (key == 12) { b2f::g_tile_cg2_min_n = value; return 0; }
  if (key == 13) { b2f::g_tile_pdl = value; return 0; }
  if (key == 14) { b2f::g_tile_wide_res = value; return 0; }
  if (key == 15) { b2f::g_tile_big_res = value; return 0; }
  if (key == 16) { b2f::g_tile_reduce = value; return 0; }
  if (key == 17) { b2f::g_match_pair = value; return 0; }
  b2f_set_error("unknown tuning key %d", key);
  return 2;
}

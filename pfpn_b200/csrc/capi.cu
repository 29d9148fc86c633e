// Status strings and ABI version of the pfpn_b200 C ABI (include/pfpn_b200.h).
#include "common.cuh"

extern "C" int pfpn_abi_version(void) { return PFPN_ABI_VERSION; }

extern "C" const char* pfpn_status_string(int status) {
  switch (status) {
    case PFPN_OK: return "ok";
    case PFPN_ERR_ARG: return "pfpn: invalid argument (null pointer, bad size or enum)";
    case PFPN_ERR_ALIGN: return "pfpn: pointer is not 16-byte aligned";
    case PFPN_ERR_UNSUPPORTED: return "pfpn: shape outside the compiled kernel instantiations";
    case PFPN_ERR_WORKSPACE: return "pfpn: workspace missing or too small";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "pfpn: unknown status";
}

// The ctypes mirror (pfpn_b200/_cabi.py: HeadArgs) assumes this layout.
#include <cstddef>
static_assert(sizeof(pfpn_head_args) == 176, "pfpn_head_args layout changed: update _cabi.HeadArgs");
static_assert(offsetof(pfpn_head_args, g_ent) == 48 && offsetof(pfpn_head_args, adv) == 56, "layout");
static_assert(offsetof(pfpn_head_args, eps_clip) == 80 && offsetof(pfpn_head_args, lp) == 88, "layout");
static_assert(offsetof(pfpn_head_args, loss) == 144 && offsetof(pfpn_head_args, B) == 152, "layout");
static_assert(sizeof(pfpn_sample_args) == 88, "pfpn_sample_args layout changed: update _cabi.SampleArgs");
static_assert(sizeof(pfpn_rsample_args) == 144 && offsetof(pfpn_rsample_args, offset_dev) == 136, "pfpn_rsample_args layout changed: update _cabi.RSampleArgs");
static_assert(sizeof(pfpn_resample_args) == 200, "pfpn_resample_args layout changed: update _cabi.ResampleArgs");
static_assert(offsetof(pfpn_resample_args, seed) == 160 && offsetof(pfpn_resample_args, threshold) == 176, "layout");
static_assert(sizeof(pfpn_head_push) == 176 && offsetof(pfpn_head_push, ticket) == 128 && offsetof(pfpn_head_push, value) == 140 &&
                  offsetof(pfpn_head_push, consume_rows) == 152 && offsetof(pfpn_head_push, consume_scale) == 168,
              "pfpn_head_push layout changed: update _cabi.HeadPush");
static_assert(sizeof(pfpn_sac_head_args) == 152 && offsetof(pfpn_sac_head_args, offset_dev) == 144, "pfpn_sac_head_args layout changed: update _cabi.SacHeadArgs");
static_assert(sizeof(pfpn_sync_args) == 208 && offsetof(pfpn_sync_args, counters) == 128 && offsetof(pfpn_sync_args, stage) == 160 && offsetof(pfpn_sync_args, rank) == 184, "pfpn_sync_args layout changed: update _cabi.SyncArgs");

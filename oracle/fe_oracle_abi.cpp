// fe_oracle_abi.cpp — the CPU ORACLE behind the SAME C symbols as the product library
// (include/fe_b200.h: fe_version, fe_params_*, fe_create, fe_set_params, fe_process_batch, fe_destroy,
// fe_last_error), so that a harness written against the C-ABI can be pointed at either shared object
// ("the CPU oracle exports the identical symbols from a second .so").
// TEST INFRASTRUCTURE ONLY — same rules as fe_oracle.cpp: nothing in the product path links or loads this.
// Every entry point forwards to the feo_* functions of fe_oracle.cpp (KD-tree mode, one thread).
#include <cstring>
#include <string>
#include <vector>

#include "../include/fe_b200.h"

extern "C" {
void feo_params_node_default(fe_params_t* p);
void feo_params_launch_playback(fe_params_t* p);
int feo_process_batch(const fe_params_t* P, const fe_point_t* points, const int64_t* scan_offsets, const double* roll_pitch,
                      int32_t n_scans, int32_t mode, int32_t n_threads, int64_t* keypoint_offsets, fe_point_t* keypoints,
                      float* descriptors, float* edge_margin, int64_t cap_kp, int64_t* n_kp_total);
}

struct fe_ctx {
  fe_params_t params;
  std::vector<int64_t> kpOffsets;
  std::vector<fe_point_t> kp;
  std::vector<float> desc;
  std::string err;
};

extern "C" {

const char* fe_version(void) { return "fe_oracle 0.2 (CPU restatement behind the fe_b200.h symbols; parity unpinned)"; }
void fe_params_node_default(fe_params_t* p) { feo_params_node_default(p); }
void fe_params_launch_playback(fe_params_t* p) { feo_params_launch_playback(p); }
int fe_device_count(void) { return 0; }

int fe_create(int device, const fe_params_t* params, const fe_limits_t* limits, fe_ctx_t** out) {
  (void)device; (void)limits;
  if (!out || !params) return FE_ERR_INVALID;
  fe_ctx* c = new fe_ctx();
  c->params = *params;
  *out = c;
  return FE_OK;
}

int fe_set_params(fe_ctx_t* ctx, const fe_params_t* params) {
  if (!ctx || !params) return FE_ERR_INVALID;
  ctx->params = *params;
  return FE_OK;
}

void fe_destroy(fe_ctx_t* ctx) { delete ctx; }
const char* fe_last_error(const fe_ctx_t* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fe_process_batch(fe_ctx_t* ctx, const fe_point_t* points, const int64_t* scan_offsets, const double* roll_pitch,
                     int32_t n_scans, fe_batch_result_t* out) {
  if (!ctx || !out || n_scans < 0 || (n_scans > 0 && (!scan_offsets || !roll_pitch))) return FE_ERR_INVALID;
  ctx->err.clear();
  ctx->kpOffsets.assign((size_t)n_scans + 1, 0);
  int64_t total = 0;
  int st = feo_process_batch(&ctx->params, points, scan_offsets, roll_pitch, n_scans, 1, 1, ctx->kpOffsets.data(), nullptr, nullptr,
                             nullptr, 0, &total);
  if (st != FE_OK) { ctx->err = "feo_process_batch (count pass) failed"; return st; }
  ctx->kp.resize((size_t)total);
  const bool desc = ctx->params.estimate_descriptors != 0;
  ctx->desc.assign(desc ? (size_t)total * FE_DESC_LEN : 0, 0.0f);
  if (total > 0) {
    st = feo_process_batch(&ctx->params, points, scan_offsets, roll_pitch, n_scans, 1, 1, ctx->kpOffsets.data(), ctx->kp.data(),
                           desc ? ctx->desc.data() : nullptr, nullptr, total, &total);
    if (st != FE_OK) { ctx->err = "feo_process_batch failed"; return st; }
  }
  out->n_scans = n_scans;
  out->n_keypoints = total;
  out->keypoint_offsets = ctx->kpOffsets.data();
  out->keypoints = ctx->kp.data();
  out->descriptors = desc ? ctx->desc.data() : nullptr;
  out->on_device = 0;
  out->gpu_launches = 0;
  return FE_OK;
}

}  // extern "C"

// fe_api.cu — host side of libfe_b200.so: context, parameter derivation, sub-batch pipeline and
// the C-ABI of include/fe_b200.h.  No CPU fallback anywhere: every entry point that computes
// launches the kernels of fe_kernels.cuh.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "fe_kernels.cuh"

using namespace fe;

#define FE_VERSION "fe_b200 0.1 (sm_100a)"

namespace {

// Small sub-batches (a single scan per callback is how the reference runs, src:72) are launch-bound: ~17 dependent
// launches.  Their whole chain — staging copies, kernels, result copies — is captured once per shape into a CUDA
// graph and replayed (fe_process_batch, sub-batches of at most GRAPH_MAX_SCANS scans).
constexpr int GRAPH_MAX_SCANS = 16;
static_assert(GRAPH_MAX_SCANS <= 32, "k_kp_offsets_gather_small gives every scan a warp of one 1024-thread block");
struct GraphKey {
  int nscans, nch, by, k1flags, stride, xo, yo, zo;
  int desc, wantKc, bnd, epoch, lean;
  const void* pts;  // device-resident input (fe_process_batch_device): the pointer is part of the captured launches
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t exec = nullptr;  // null: the shape has been seen once (run eagerly, so that every lazy init is done)
  int64_t launches = 0;
};

struct Slot {
  cudaStream_t stream = nullptr;
  // lean chain (see process_batch_host): what finalize_subbatch needs to run the sub-batch again with every fallback kernel
  bool lean = false;
  const float4* rrPts = nullptr;
  int rrNch = 0, rrK1flags = 0, rrStride = 0, rrXo = 0, rrYo = 0, rrZo = 0;
  bool rrDesc = false, rrWantKc = false, rrRaw = false;
  std::vector<GraphEntry> graphs;
  int64_t capPts = 0;
  int capScans = 0, capChunks = 0;
  int64_t capKp = 0;
  int capKf = 0, capKc = 0;
  // device
  float4 *d_pts = nullptr, *d_surf = nullptr, *d_crop = nullptr, *d_sorted = nullptr, *d_full = nullptr;
  float4* d_ringPts = nullptr;  // K2: the crop survivors bucketed by ring (CSR like the input)
  unsigned *d_cropMeta = nullptr, *d_keyA = nullptr, *d_keyB = nullptr, *d_valA = nullptr, *d_valB = nullptr, *d_sortedKey = nullptr;
  int* d_rho = nullptr;
  long long* d_scan_off = nullptr;
  int *d_chunk_off = nullptr, *d_surfCnt = nullptr, *d_cropCnt = nullptr;
  float* d_rot = nullptr;
  int *d_kfBase = nullptr, *d_kfCnt = nullptr, *d_kcBase = nullptr, *d_kcCnt = nullptr;
  int *d_kpBase = nullptr, *d_kpCnt = nullptr, *d_kpOff = nullptr, *d_kpScan = nullptr, *d_kpNbr = nullptr;
  int *d_kpNbrOff = nullptr, *d_kpRank = nullptr;  // K4b -> K4d: slice of the neighbour pool (d_keyA), RNG rank
  int *d_kpListM = nullptr, *d_kpListL = nullptr;  // keypoints of K4d's medium / large instantiation
  int *d_rowStart = nullptr, *d_surfN = nullptr, *d_perScan = nullptr, *d_outOff = nullptr;
  int *d_perScan2 = nullptr, *d_outOff2 = nullptr;   // cloud outputs: ~keypoint_cloud counts / offsets
  float4* d_gather2 = nullptr;
  int *h_cloudOff = nullptr, *h_kcOff = nullptr;     // pinned: per-scan offsets of the two cloud outputs of the sub-batch
  unsigned char* d_gridHdr = nullptr;  // per-scan index of the halo cell grid (bit map, word prefix, slot table)
  int* d_tabOk = nullptr;
  int64_t capGridHdr = 0;
  int *d_ringBase = nullptr, *d_ovfRuns = nullptr, *d_scanFlag = nullptr;  // K2: ring segment offsets per scan, rings for the wide run kernel
  int *d_ovfRings = nullptr, *d_ovfRings2 = nullptr, *d_ovfMerge = nullptr, *d_ovfMerge2 = nullptr, *d_ovfSurf = nullptr;
  unsigned char* d_slabs = nullptr;
  int64_t capRowStart = 0;
  float4 *d_kfPool = nullptr, *d_kcPool = nullptr, *d_kpPool = nullptr, *d_kpOut = nullptr, *d_gather = nullptr;
  float* d_desc = nullptr;
  DevCounters* d_ctr = nullptr;
  unsigned long long* d_bnd = nullptr;  // tolerance-boundary report: [scan][4] pair counts
  unsigned long long* h_bnd = nullptr;
  // pinned host mirrors
  long long* h_scan_off = nullptr;
  int* h_chunk_off = nullptr;
  int4* d_chunkTab = nullptr;  // k_chunk_table, per chunk: scan, (chunk of scan << 12) | points, first point lo, hi
  float* h_rot = nullptr;
  DevCounters* h_ctr = nullptr;
  int* h_kpOff = nullptr;
  int* h_perScan = nullptr;
  cudaEvent_t evDone = nullptr, evT0 = nullptr, evT1 = nullptr;
  // K4a depends on K1 only: it runs on a side stream next to K2/K3 (fork after K1, join before K4b)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  bool sideK4a = false;  // the pipeline in flight uses the side stream
  int lastNch = 0;
  int64_t maxScanPts = 0;  // largest scan of the staged sub-batch
  // bookkeeping of the sub-batch in flight
  int nscans = 0;
  int64_t npts = 0;
  int firstScan = 0;
  bool busy = false;
  // stage timing
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> evName;
  int nev = 0;
};

}  // namespace

struct fe_ctx {
  int device = 0;
  fe_params_t params;
  DevParams dp;
  fe_limits_t lim;
  Slot slot[2];
  float* d_lut = nullptr;
  float2* d_axes = nullptr;
  int axesCap = 0;
  bool cloudOutputs = false;
  int leanHold = 0;           // sub-batches for which the lean chain stays off after one had to be run again
  int64_t leanReruns = 0;
  bool stageTiming = false;   // serialise the stages and time each with CUDA events (fe_enable_stage_timing)
  bool gridClustering = false;  // fe_debug_force_grid_clustering: K2 through the grid-based kernels only
  bool recordOutput = false;  // descriptors leave as FE_RECORD_FLOATS-float PointDescriptor records
  int numSms = 148;           // cudaDevAttrMultiProcessorCount of the context's device (grids are sized in resident waves)
  int epoch = 0;              // bumped by every setting that changes what a captured graph would do
  bool useGraphs = true;      // fe_debug_enable_graphs
  int64_t graphReplays = 0;
  int angleLibm = 0;          // fe_set_angle_libm
  double bndEps = 0.0;        // fe_enable_boundary_report: > 0 = count the pairs within bndEps of every radius
  std::vector<int64_t> bndCounts;  // [scan][4] of the last batch call
  // results (host, pinned, grown on demand)
  std::vector<int64_t> kpOffsets;
  fe_point_t* h_kp = nullptr;
  float* h_desc = nullptr;
  int64_t capResKp = 0, capResDesc = 0;
  // optional cloud outputs of the last host batch call (fe_enable_cloud_outputs): CSR by scan, pinned
  std::vector<int64_t> cloudOff, kcOff;
  fe_point_t *h_cloud = nullptr, *h_kc = nullptr;
  int64_t capCloud = 0, capKcRes = 0, cloudRun = 0, kcRun = 0;
  // stage times of the last device call
  std::vector<const char*> stName;
  std::vector<float> stMs;
  int64_t launches = 0;
  std::string err;
};

namespace {

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      char b_[512];                                                                    \
      snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = b_;                                                                   \
      return FE_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

int fail(fe_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

// ---- parameter derivation (host, float/double exactly as PCL narrows them) ------------------

float radius_sq_as_flann_sees_it(double radius) { return (float)(radius * radius); }

// pcl::ShapeContext3DEstimation::initCompute (3dsc.hpp): bin edges and the 1/cbrt(volume) table
void shape_context_tables(double R, double rmin, float radii[16], float theta[12], float phi[13], float* lut) {
  const int NA = 12, NE = 11, NR = 15;
  const float az_step = 360.0f / (float)NA, el_step = 180.0f / (float)NE;
  for (int j = 0; j <= NR; j++)
    radii[j] = (float)exp(log(rmin) + (((float)j / (float)NR) * log(R / rmin)));
  for (int k = 0; k <= NE; k++) theta[k] = (float)k * el_step;
  for (int l = 0; l <= NA; l++) phi[l] = (float)l * az_step;
  const float d2r = 0.017453293f;
  const float dphi = phi[1] * d2r - phi[0] * d2r;
  const float third = 1.0f / 3.0f;
  for (int j = 0; j < NR; j++) {
    const float dr = (radii[j + 1] * radii[j + 1] * radii[j + 1] / 3.0f) - (radii[j] * radii[j] * radii[j] / 3.0f);
    for (int k = 0; k < NE; k++) {
      const float dth = cosf(theta[k] * d2r) - cosf(theta[k + 1] * d2r);
      const float V = dphi * dth * dr;
      for (int l = 0; l < NA; l++) lut[l * NE * NR + k * NR + j] = 1.0f / powf(V, third);
    }
  }
}

// Eigen 3.2 AngleAxisf(pitch, Y) * AngleAxisf(roll, X) -> rotation matrix (row-major 3x3)
void leveling_matrix(double roll, double pitch, float m[9]) {
  const float hp = 0.5f * (float)pitch, hr = 0.5f * (float)roll;
  // quaternions (w, x, y, z)
  const float aw = cosf(hp), ax = 0.0f * sinf(hp), ay = 1.0f * sinf(hp), az = 0.0f * sinf(hp);
  const float bw = cosf(hr), bx = 1.0f * sinf(hr), by = 0.0f * sinf(hr), bz = 0.0f * sinf(hr);
  const float w = aw * bw - ax * bx - ay * by - az * bz;
  const float x = aw * bx + ax * bw + ay * bz - az * by;
  const float y = aw * by + ay * bw + az * bx - ax * bz;
  const float z = aw * bz + az * bw + ax * by - ay * bx;
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  float R[9];
  R[0] = 1.0f - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
  R[3] = txy + twz;          R[4] = 1.0f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = 1.0f - (txx + tyy);
  // Affine3f::Identity().rotate(q): Identity.linear() * R, 3-term sums as a0 + (a1 + a2)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const float i0 = (i == 0) ? 1.0f : 0.0f, i1 = (i == 1) ? 1.0f : 0.0f, i2 = (i == 2) ? 1.0f : 0.0f;
      m[i * 3 + j] = i0 * R[j] + (i1 * R[3 + j] + i2 * R[6 + j]);
    }
}

int bits_for_host(int v) { int b = 0; while (v > 0) { b++; v >>= 1; } return b; }

// surface keep-box and 2-D grid for keypoints known to lie inside [x0,x1]x[y0,y1]x[z0,z1]
void surface_grid(DevParams& dp, double R, float x0, float x1, float y0, float y1, float z0, float z1) {
  const double rrho = R / 5.0;
  const float reach = (float)((R + rrho) * 1.001 + 1e-3);
  dp.sx0 = x0 - reach; dp.sx1 = x1 + reach;
  dp.sy0 = y0 - reach; dp.sy1 = y1 + reach;
  dp.sz0 = z0 - reach; dp.sz1 = z1 + reach;
  float cell = (float)(rrho * 1.001);
  if (!(cell > 1e-6f)) cell = 1e-6f;
  const float ex = dp.sx1 - dp.sx0, ey = dp.sy1 - dp.sy0;
  const float maxdim = 2047.0f;
  if (ex / cell > maxdim) cell = ex / maxdim;
  if (ey / cell > maxdim) cell = ey / maxdim;
  dp.sg_inv = 1.0f / cell;
  dp.sg_nx = std::max(1, std::min(2048, (int)floorf(ex * dp.sg_inv) + 1));
  dp.sg_ny = std::max(1, std::min(2048, (int)floorf(ey * dp.sg_inv) + 1));
  dp.sg_bx = std::max(1, bits_for_host(dp.sg_nx - 1));
  dp.Rpad = (float)(R * 1.0001 + 1e-6);
  dp.rhopad = (float)(rrho * 1.0001 + 1e-6);
  dp.halopad = (float)((R + rrho) * 1.0002 + 4e-6);
  dp.zs0 = z0 - (float)(2.0 * rrho);  // z slabs of the cell grid: uniform over the keypoints' z range grown by 2 R/5, clamped outside
  dp.zs_inv_range = 1.0f / std::max((z1 + (float)(2.0 * rrho)) - dp.zs0, 1e-3f);
}

int derive_params(fe_ctx* ctx, const fe_params_t& p) {
  if (!(p.descriptor_radius > 0.0) || !(p.cluster_tolerance > 0.0) || !(p.cluster_radius_threshold > 0.0))
    return fail(ctx, FE_ERR_INVALID, "cluster_tolerance, cluster_radius_threshold and descriptor_radius must be > 0");
  DevParams& dp = ctx->dp;
  memset(&dp, 0, sizeof dp);
  dp.xmin = (float)p.x_min; dp.xmax = (float)p.x_max;
  dp.ymin = (float)p.y_min; dp.ymax = (float)p.y_max;
  dp.zmin = (float)p.z_min; dp.zmax = (float)p.z_max;
  dp.tol_f = (float)p.cluster_tolerance;
  dp.r2f_cluster = radius_sq_as_flann_sees_it((double)dp.tol_f);
  dp.min_count = p.cluster_min_count;
  dp.max_count = p.cluster_max_count;
  dp.two_radius_threshold = 2 * p.cluster_radius_threshold;
  dp.radius_threshold = p.cluster_radius_threshold;
  dp.merge_tol_f = (float)p.cluster_radius_threshold;
  dp.r2f_merge = radius_sq_as_flann_sees_it((double)dp.merge_tol_f);
  dp.min_channels = p.number_detection_channels;
  dp.R2f = radius_sq_as_flann_sees_it(p.descriptor_radius);
  dp.rho2f = radius_sq_as_flann_sees_it(p.descriptor_radius / 5.0);
  dp.estimate_descriptors = p.estimate_descriptors;
  dp.angle_libm = ctx->angleLibm;
  std::vector<float> lut(FE_DESC_LEN);
  shape_context_tables(p.descriptor_radius, p.descriptor_radius / 10.0, dp.radii, dp.theta, dp.phi, lut.data());
  surface_grid(dp, p.descriptor_radius, dp.xmin, dp.xmax, dp.ymin, dp.ymax, dp.zmin, dp.zmax);
  CK(cudaMemcpy(ctx->d_lut, lut.data(), FE_DESC_LEN * sizeof(float), cudaMemcpyHostToDevice));
  ctx->params = p;
  return FE_OK;
}

// ---- slot allocation ---------------------------------------------------------------------------

template <class T>
cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }
template <class T>
cudaError_t halloc(T** p, size_t n) { return cudaHostAlloc((void**)p, std::max<size_t>(n, 1) * sizeof(T), cudaHostAllocDefault); }

void free_slot(Slot& s) {
  void* dv[] = {s.d_ringPts, s.d_pts, s.d_surf, s.d_crop, s.d_sorted, s.d_full, s.d_cropMeta, s.d_keyA, s.d_keyB, s.d_valA, s.d_valB,
                s.d_sortedKey, s.d_rho, s.d_scan_off, s.d_chunk_off, s.d_surfCnt, s.d_cropCnt, s.d_rot, s.d_kfBase,
                s.d_kfCnt, s.d_kcBase, s.d_kcCnt, s.d_kpBase, s.d_kpCnt, s.d_kpOff, s.d_kpScan, s.d_kpNbr, s.d_kpNbrOff, s.d_kpRank, s.d_kpListM, s.d_kpListL, s.d_rowStart,
                s.d_surfN, s.d_perScan, s.d_outOff, s.d_ovfRings, s.d_ovfRings2, s.d_ovfMerge, s.d_ovfMerge2, s.d_ovfSurf, s.d_slabs, s.d_gridHdr, s.d_tabOk, s.d_kfPool, s.d_kcPool, s.d_kpPool, s.d_kpOut, s.d_gather, s.d_desc, s.d_ctr, s.d_bnd, s.d_ringBase, s.d_ovfRuns, s.d_scanFlag, s.d_perScan2, s.d_outOff2, s.d_gather2, s.d_chunkTab};
  for (void* p : dv) if (p) cudaFree(p);
  void* hv[] = {s.h_scan_off, s.h_chunk_off, s.h_rot, s.h_ctr, s.h_kpOff, s.h_perScan, s.h_bnd, s.h_cloudOff, s.h_kcOff};
  for (void* p : hv) if (p) cudaFreeHost(p);
  for (cudaEvent_t e : s.ev) cudaEventDestroy(e);
  for (GraphEntry& g : s.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (s.evDone) cudaEventDestroy(s.evDone);
  if (s.evT0) cudaEventDestroy(s.evT0);
  if (s.evT1) cudaEventDestroy(s.evT1);
  for (cudaEvent_t e : {s.evFork, s.evJoin}) if (e) cudaEventDestroy(e);
  if (s.stream2) cudaStreamDestroy(s.stream2);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Slot();
}

int ensure_slot(fe_ctx* ctx, Slot& s, bool ownPoints) {
  if (s.stream) {
    if (ownPoints && !s.d_pts) CK(dalloc(&s.d_pts, (size_t)s.capPts));
    return FE_OK;
  }
  const fe_limits_t& L = ctx->lim;
  s.capPts = L.max_points_per_call;
  s.capScans = L.max_scans_per_call;
  s.capChunks = (int)(s.capPts / CH) + s.capScans + 1;
  s.capKp = L.max_keypoints_per_call;
  s.capKf = (int)L.max_ring_clusters_per_call;
  s.capKc = 0;
  CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&s.evFork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.evJoin, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.evDone, cudaEventDisableTiming));
  CK(cudaEventCreate(&s.evT0));
  CK(cudaEventCreate(&s.evT1));
  const size_t np = (size_t)s.capPts, ns = (size_t)s.capScans;
  if (ownPoints) CK(dalloc(&s.d_pts, np));
  CK(dalloc(&s.d_surf, np)); CK(dalloc(&s.d_crop, np)); CK(dalloc(&s.d_sorted, np)); CK(dalloc(&s.d_ringPts, np));
  CK(dalloc(&s.d_cropMeta, np)); CK(dalloc(&s.d_keyA, np)); CK(dalloc(&s.d_keyB, np));
  CK(dalloc(&s.d_valA, np)); CK(dalloc(&s.d_valB, np)); CK(dalloc(&s.d_sortedKey, np));
  CK(dalloc(&s.d_rho, np));
  CK(dalloc(&s.d_scan_off, ns + 1)); CK(dalloc(&s.d_chunk_off, ns + 1));
  CK(dalloc(&s.d_surfCnt, (size_t)s.capChunks)); CK(dalloc(&s.d_cropCnt, (size_t)s.capChunks));
  CK(dalloc(&s.d_chunkTab, (size_t)s.capChunks));
  CK(dalloc(&s.d_rot, ns * 9));
  CK(dalloc(&s.d_kfBase, ns * 16)); CK(dalloc(&s.d_kfCnt, ns * 16));
  CK(dalloc(&s.d_kcBase, ns * 16)); CK(dalloc(&s.d_kcCnt, ns * 16));
  CK(dalloc(&s.d_kpBase, ns)); CK(dalloc(&s.d_kpCnt, ns)); CK(dalloc(&s.d_kpOff, ns + 1));
  CK(dalloc(&s.d_kpScan, (size_t)s.capKp)); CK(dalloc(&s.d_kpNbr, (size_t)s.capKp));
  CK(dalloc(&s.d_kpNbrOff, (size_t)s.capKp)); CK(dalloc(&s.d_kpRank, (size_t)s.capKp));
  CK(dalloc(&s.d_kpListM, (size_t)s.capKp)); CK(dalloc(&s.d_kpListL, (size_t)s.capKp));
  CK(dalloc(&s.d_surfN, ns)); CK(dalloc(&s.d_perScan, ns)); CK(dalloc(&s.d_outOff, ns + 1)); CK(dalloc(&s.d_tabOk, ns));
  CK(dalloc(&s.d_ringBase, ns * 17)); CK(dalloc(&s.d_ovfRuns, ns * 16)); CK(dalloc(&s.d_scanFlag, ns));
  CK(dalloc(&s.d_ovfRings, ns)); CK(dalloc(&s.d_ovfRings2, ns)); CK(dalloc(&s.d_ovfMerge, ns)); CK(dalloc(&s.d_ovfMerge2, ns));
  CK(dalloc(&s.d_ovfSurf, ns));
  CK(dalloc(&s.d_slabs, (size_t)NGLOBAL * cluster_slab_bytes(ECAP_G)));
  CK(dalloc(&s.d_kfPool, (size_t)s.capKf)); CK(dalloc(&s.d_kpPool, (size_t)s.capKp)); CK(dalloc(&s.d_kpOut, (size_t)s.capKp));
  CK(dalloc(&s.d_desc, (size_t)s.capKp * FE_RECORD_FLOATS));  // room for either output layout
  CK(dalloc(&s.d_ctr, 1));
  CK(halloc(&s.h_scan_off, ns + 1)); CK(halloc(&s.h_chunk_off, ns + 1)); CK(halloc(&s.h_rot, ns * 9));
  CK(halloc(&s.h_ctr, 1)); CK(halloc(&s.h_kpOff, ns + 1)); CK(halloc(&s.h_perScan, ns + 1));
  s.ev.resize(24);
  for (auto& e : s.ev) CK(cudaEventCreate(&e));
  s.evName.assign(24, "");
  return FE_OK;
}

void mark(fe_ctx* ctx, Slot& s, const char* name) {
  if (ctx->stageTiming && s.nev < (int)s.ev.size()) {
    cudaEventRecord(s.ev[s.nev], s.stream);
    s.evName[s.nev] = name;
    s.nev++;
  }
}

std::string err_bits(int e) {
  std::string m;
  if (e & ERR_RING_CAP) m += "one ring of a scan holds more cluster entries than the shared-memory capacity; ";
  if (e & ERR_KF_POOL) m += "ring-centroid pool exhausted (raise max_ring_clusters_per_call); ";
  if (e & ERR_KP_POOL) m += "keypoint pool exhausted (raise max_keypoints_per_call); ";
  if (e & ERR_MERGE_CAP) m += "a scan has more ring centroids than the merge capacity; ";
  if (e & ERR_AXIS_CAP) m += "a scan has more keypoints than precomputed 3DSC axes; ";
  if (e & ERR_CHUNKS) m += "a scan has more points than the per-scan kernels index; ";
  if (e & ERR_KC_POOL) m += "keypoint_cloud pool exhausted; ";
  return m;
}

// Host-side staging of the small per-scan arrays of a sub-batch.  offs are absolute offsets of
// the caller's array; the device sees offsets relative to the first point of the sub-batch.
// the device copies of the staged arrays (from the pinned mirrors: the same addresses every call, so the copies can
// live inside a captured graph)
int stage_scans_copy(fe_ctx* ctx, Slot& s, int nscans, bool deferRot) {
  CK(cudaMemcpyAsync(s.d_scan_off, s.h_scan_off, (nscans + 1) * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
  CK(cudaMemcpyAsync(s.d_chunk_off, s.h_chunk_off, (nscans + 1) * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  if (s.h_chunk_off[nscans] > 0) {  // K1's chunk -> scan table, built where the offsets just landed
    k_chunk_table<<<(nscans + 7) / 8, 256, 0, s.stream>>>(s.d_scan_off, s.d_chunk_off, nscans, s.d_chunkTab);
    ctx->launches++;
  }
  if (!deferRot) CK(cudaMemcpyAsync(s.d_rot, s.h_rot, (size_t)nscans * 9 * sizeof(float), cudaMemcpyHostToDevice, s.stream));
  return FE_OK;
}

int stage_scans(fe_ctx* ctx, Slot& s, const int64_t* offs, const double* rp, int nscans, int64_t* nptsOut, int* nchOut,
                bool deferRot = false, bool copy = true) {
  // the pinned mirrors hold capScans(+1) entries: check before the first write
  if (nscans > s.capScans) return fail(ctx, FE_ERR_CAPACITY, "sub-batch exceeds max_scans_per_call");
  const int64_t o0 = offs[0];
  if (o0 < 0) return fail(ctx, FE_ERR_INVALID, "scan_offsets must not be negative");
  if (offs[nscans] - o0 > s.capPts) return fail(ctx, FE_ERR_CAPACITY, "sub-batch exceeds max_points_per_call");
  int nch = 0;
  s.maxScanPts = 0;
  for (int i = 0; i < nscans; i++) {
    const int64_t n = offs[i + 1] - offs[i];
    if (n < 0) return fail(ctx, FE_ERR_INVALID, "scan_offsets must be non-decreasing");
    s.maxScanPts = std::max(s.maxScanPts, n);
    s.h_scan_off[i] = (long long)(offs[i] - o0);
    s.h_chunk_off[i] = nch;
    const int64_t nc = (n + CH - 1) / CH;
    if (nc > (int64_t)s.capChunks - nch) return fail(ctx, FE_ERR_CAPACITY, "sub-batch exceeds the chunk capacity");
    if (nc >= (1 << 20)) return fail(ctx, FE_ERR_CAPACITY, "a scan of more than 2^31 points");  // k_chunk_table packs (chunk << 12) | points
    nch += (int)nc;
    if (deferRot) continue;  // enqueue_pipeline computes the matrices range by range (lateRp)
    if (rp) leveling_matrix(rp[2 * i], rp[2 * i + 1], s.h_rot + 9 * i);
    else { float* m = s.h_rot + 9 * i; for (int k = 0; k < 9; k++) m[k] = (k % 4 == 0) ? 1.0f : 0.0f; }
  }
  s.h_scan_off[nscans] = (long long)(offs[nscans] - o0);
  s.h_chunk_off[nscans] = nch;
  *nptsOut = offs[nscans] - o0;
  *nchOut = nch;
  if (nch > s.capChunks) return fail(ctx, FE_ERR_CAPACITY, "sub-batch exceeds the chunk capacity");
  if (copy) return stage_scans_copy(ctx, s, nscans, deferRot);
  return FE_OK;
}

int ensure_rowstart(fe_ctx* ctx, Slot& s, int nscans) {
  const int64_t need = (int64_t)std::max(nscans, s.capScans) * (ctx->dp.sg_ny + 1);
  if (need > s.capRowStart) {
    if (s.d_rowStart) { CK(cudaStreamSynchronize(s.stream)); CK(cudaFree(s.d_rowStart)); s.d_rowStart = nullptr; }
    CK(dalloc(&s.d_rowStart, (size_t)need));
    s.capRowStart = need;
  }
  const int64_t ncells = (int64_t)ctx->dp.sg_nx * ctx->dp.sg_ny;
  if (ncells <= SURF_MAX_CELLS) {
    const int64_t needT = (int64_t)std::max(nscans, s.capScans) * grid_hdr_bytes((int)ncells);
    if (needT > s.capGridHdr) {
      if (s.d_gridHdr) { CK(cudaStreamSynchronize(s.stream)); CK(cudaFree(s.d_gridHdr)); s.d_gridHdr = nullptr; }
      CK(dalloc(&s.d_gridHdr, (size_t)needT));
      s.capGridHdr = needT;
    }
  }
  return FE_OK;
}

SurfIndex surf_index(const fe_ctx* ctx, const Slot& s) {
  const int64_t ncells = (int64_t)ctx->dp.sg_nx * ctx->dp.sg_ny;
  SurfIndex X;
  X.sortedKey = s.d_sortedKey;
  X.rowStart = s.d_rowStart;
  X.gridHdr = (ncells <= SURF_MAX_CELLS) ? s.d_gridHdr : nullptr;
  X.gridStride = grid_hdr_bytes((int)ncells);
  X.tabOk = s.d_tabOk;
  return X;
}

int ensure_kc(fe_ctx* ctx, Slot& s) {
  if (s.capKc == 0) {
    s.capKc = (int)std::min<int64_t>(s.capPts, 1 << 26);
    CK(dalloc(&s.d_kcPool, (size_t)s.capKc));
    CK(dalloc(&s.d_gather, (size_t)s.capPts));
    CK(dalloc(&s.d_gather2, (size_t)s.capKc));
    CK(dalloc(&s.d_perScan2, (size_t)s.capScans));
    CK(dalloc(&s.d_outOff2, (size_t)s.capScans + 1));
    CK(halloc(&s.h_cloudOff, (size_t)s.capScans + 1));
    CK(halloc(&s.h_kcOff, (size_t)s.capScans + 1));
  }
  return FE_OK;
}

const size_t kClusterSmem = cluster_smem_bytes(ECAP, NTF);
const size_t kClusterSmemL = cluster_smem_bytes(ECAP_L, NT2);
const size_t kClusterSmemL2 = cluster_smem_bytes(ECAP_L, NTL);
const size_t kClusterSmemM = cluster_smem_bytes(ECAP_M, NTM);

int set_kernel_attrs(fe_ctx* ctx) {
  CK(cudaFuncSetAttribute(k_cluster_rings<ECAP, NTF, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem));
  CK(cudaFuncSetAttribute(k_ring_runs<NW_RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RingRunsSmT<RW>)));
  CK(cudaFuncSetAttribute(k_ring_runs<NW_RR_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RingRunsSmT<RW, NW_RR_SMALL>)));
  CK(cudaFuncSetAttribute(k_ring_runs_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NW_RR * sizeof(RunBufT<RW2>))));
  CK(cudaFuncSetAttribute(k_cluster_rings<ECAP_L, NTL, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemL2));
  CK(cudaFuncSetAttribute(k_merge_keypoints<ECAP_M, NTM, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemM));
  CK(cudaFuncSetAttribute(k_merge_keypoints<ECAP_L, NT2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemL));
  CK(cudaFuncSetAttribute(k_extract_clusters_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemL));
  CK(cudaFuncSetAttribute(k_desc_hist_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)desc_warp_smem_bytes()));
  CK(cudaFuncSetAttribute(k_desc_hist<256, DCAP, DW_CAP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)desc_smem_bytes(DCAP, 256)));
  CK(cudaFuncSetAttribute(k_desc_hist<512, DCAP_M, DCAP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)desc_smem_bytes(DCAP_M, 512)));
  CK(cudaFuncSetAttribute(k_desc_hist<512, DCAP_L, DCAP_M, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)desc_smem_bytes(DCAP_L, 512)));
  CK(cudaFuncSetAttribute(k_surface_grid_cells<NT_SURF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)surf_cells_smem_bytes(SURF_MAX_CELLS)));
  return FE_OK;
}

// K4d: three instantiations split the keypoints by neighbour count (smaller footprint = more blocks / SM)
void launch_desc_hist(fe_ctx* ctx, Slot& s, int nscans, const DevParams& P, int gridKp, bool records, bool lean = false) {
  const int descStride = records ? FE_RECORD_FLOATS : FE_DESC_LEN, descOff = records ? 5 : 0;
#define FE_DESC_ARGS s.d_kpOut, s.d_kpScan, s.d_kpOff, nscans, s.d_kpNbr, s.d_sorted, surf_index(ctx, s), s.d_scan_off, P, \
                     s.d_rho, ctx->d_lut, ctx->d_axes, ctx->axesCap, s.d_keyA, s.d_kpNbrOff, s.d_kpRank, FE_DESC_LIST, s.d_desc, descStride, descOff, s.d_ctr, warpCap
  // the blocks stride over the keypoints with equal shares: grids of exactly one resident wave
  static const int perSm = []() {  // initialised once, also when fe_multi's threads get here together
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_desc_hist<256, DCAP, DW_CAP, false, true>, 256, desc_smem_bytes(DCAP, 256)) != cudaSuccess) v = 4;
    return v;
  }();
  // A handful of scans (the reference's one scan per callback) has too few keypoints to fill the GPU with warps:
  // there the same algorithm runs with a block per keypoint (k_desc_hist_small), several times shorter per keypoint.
  const int warpCap = DW_CAP;
  if (nscans <= GRAPH_MAX_SCANS) {
    k_desc_hist_small<<<std::min(gridKp, ctx->numSms * 4), DW_CAP, 0, s.stream>>>(
        s.d_kpOut, s.d_kpScan, s.d_kpOff, nscans, s.d_kpNbr, s.d_sorted, s.d_scan_off, P, s.d_rho, ctx->d_lut, ctx->d_axes, ctx->axesCap,
        s.d_keyA, s.d_kpNbrOff, s.d_kpRank, s.d_desc, descStride, descOff, s.d_ctr);
    ctx->launches++;
  } else {  // keypoints with at most DW_CAP listed neighbours (and the empty ones): a warp each
    static const int perSmW = []() {
      int v = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_desc_hist_warp, DW_WARPS * 32, desc_warp_smem_bytes()) != cudaSuccess) v = 2;
      return v;
    }();
    const int nsm = ctx->numSms;
    k_desc_hist_warp<<<nsm * std::max(perSmW, 1), DW_WARPS * 32, desc_warp_smem_bytes(), s.stream>>>(
        s.d_kpOut, s.d_kpScan, s.d_kpOff, nscans, s.d_kpNbr, s.d_sorted, s.d_scan_off, P, s.d_rho, ctx->d_lut, ctx->d_axes, ctx->axesCap,
        s.d_keyA, s.d_kpNbrOff, s.d_kpRank, s.d_desc, descStride, descOff, s.d_ctr);
    ctx->launches++;
  }
#define FE_DESC_LIST nullptr, nullptr
  k_desc_hist<256, DCAP, DW_CAP, false, true><<<std::min(gridKp, ctx->numSms * std::max(perSm, 1)), 256, desc_smem_bytes(DCAP, 256), s.stream>>>(FE_DESC_ARGS);
#undef FE_DESC_LIST
  ctx->launches++;
  if (!lean) {  // lean: keypoints listed for the two larger instantiations make the caller run the sub-batch again
#define FE_DESC_LIST s.d_kpListM, &s.d_ctr->n_list_m
  k_desc_hist<512, DCAP_M, DCAP, false, false><<<std::min(gridKp, ctx->numSms * 2), 512, desc_smem_bytes(DCAP_M, 512), s.stream>>>(FE_DESC_ARGS);
#undef FE_DESC_LIST
#define FE_DESC_LIST s.d_kpListL, &s.d_ctr->n_list_l
  k_desc_hist<512, DCAP_L, DCAP_M, true, false><<<std::min(gridKp, ctx->numSms), 512, desc_smem_bytes(DCAP_L, 512), s.stream>>>(FE_DESC_ARGS);
#undef FE_DESC_LIST
  ctx->launches += 2;
  }
#undef FE_DESC_ARGS
  if (records) {
    k_record_frame<<<ctx->numSms * 2, 256, 0, s.stream>>>(s.d_kpOut, s.d_kpOff, nscans, s.d_desc, descStride, FE_DESC_LEN);
    ctx->launches++;
  }
}

// K4b: mark + count + neighbour lists (the pool is d_keyA, idle once K4a is done), then the RNG ranks.
void launch_desc_mark(fe_ctx* ctx, Slot& s, int nscans, const DevParams& P, int gridKp) {
  const long long nbrCap = (long long)std::min<int64_t>(s.capPts, 0x7fffffff - MARK_LCAP);
  // the blocks stride over the keypoints with equal shares: a grid of exactly one resident wave
  static const int perSm = []() {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_desc_mark, 256, MARK_LCAP * sizeof(unsigned)) != cudaSuccess) v = 4;
    return v;
  }();
  gridKp = std::min(gridKp, ctx->numSms * std::max(perSm, 1));
  k_desc_mark<<<gridKp, 256, MARK_LCAP * sizeof(unsigned), s.stream>>>(s.d_kpOut, s.d_kpScan, s.d_kpOff, nscans, s.d_sorted,
                                                                       surf_index(ctx, s), s.d_scan_off, P, s.d_rho, s.d_kpNbr,
                                                                       s.d_keyA, nbrCap, s.d_kpNbrOff, s.d_ctr);
  k_kp_rank<<<(int)((s.capKp + 255) / 256), 256, 0, s.stream>>>(s.d_kpScan, s.d_kpOff, nscans, s.d_kpNbr, s.d_kpRank, s.d_kpListM, s.d_kpListL, s.d_ctr);
  ctx->launches += 2;
}

// K4a: counting sort by cell in shared memory for scans that fit (<= 65,535 surface points, grid <=
// SURF_MAX_CELLS cells), radix sort through global memory for the deferred rest.
void launch_surface_grid(fe_ctx* ctx, Slot& s, int nscans, const DevParams& P, cudaStream_t q) {
  const int ncells = P.sg_nx * P.sg_ny;
  if (ncells <= SURF_MAX_CELLS) {
    k_surface_grid_cells<NT_SURF><<<nscans, NT_SURF, surf_cells_smem_bytes(ncells), q>>>(
        s.d_surf, s.d_surfCnt, s.d_scan_off, s.d_chunk_off, P, s.d_kpOut, s.d_kpOff, s.d_sorted, s.d_rho,
        s.d_surfN, s.d_ctr, s.d_ovfSurf, s.d_gridHdr, grid_hdr_bytes(ncells), s.d_tabOk);
    ctx->launches++;
    // the radix kernel only takes scans the counting sort defers: more than 65,535 surface points or more halo
    // cells than counters — impossible when every scan is small and the grid has no more cells than counters
    if (s.maxScanPts <= 65535 && ncells <= GRID_TAB_CAP) return;
    k_surface_grid<<<std::min(nscans, ctx->numSms * 2), NT2, 0, q>>>(s.d_surf, s.d_surfCnt, s.d_scan_off, s.d_chunk_off, P, s.d_keyA,
                                                                   s.d_keyB, s.d_valA, s.d_valB, s.d_sorted, s.d_sortedKey,
                                                                   s.d_rowStart, s.d_surfN, s.d_ctr, s.d_ovfSurf, &s.d_ctr->ovf_surf,
                                                                   s.d_tabOk, s.d_rho);
  } else {
    k_surface_grid<<<nscans, NT2, 0, q>>>(s.d_surf, s.d_surfCnt, s.d_scan_off, s.d_chunk_off, P, s.d_keyA, s.d_keyB, s.d_valA,
                                                 s.d_valB, s.d_sorted, s.d_sortedKey, s.d_rowStart, s.d_surfN, s.d_ctr, nullptr, nullptr,
                                                 s.d_tabOk, s.d_rho);
  }
  ctx->launches++;
}

// blocks per scan of K4c: about one per 8k points of an average scan (a dense scan's halo is tens of tiles)
int density_blocks_per_scan(int nscans, int64_t npts) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(16, npts / std::max(nscans, 1) / 8192 + 1));
}

void launch_density(fe_ctx* ctx, Slot& s, int nscans, const DevParams& P, int64_t npts) {
  if (nscans <= 0) return;
  const int by = density_blocks_per_scan(nscans, npts);
  k_density<<<dim3((unsigned)nscans, (unsigned)by), 256, 0, s.stream>>>(s.d_sorted, surf_index(ctx, s), s.d_scan_off, P, s.d_surfN, s.d_rho, s.d_ctr->dens_work);
  ctx->launches++;
}

// K2 + K3 for `nscans` scans: the 2-blocks-per-SM instantiation first, then the large one over
// whatever scans it deferred (an immediate exit when there are none).
void mark(fe_ctx* ctx, Slot& s, const char* name);

void launch_clustering(fe_ctx* ctx, Slot& s, int nscans, bool singleRing, bool wantKc, bool merge, bool marks = false, bool lean = false) {
  // Every stage is a chain of instantiations: the fast one takes all scans and defers those that do
  // not fit its shared memory to a list; the large shared-memory one takes that list; what does not
  // fit there either goes to the instantiation whose per-entry arrays live in global memory.
  const DevParams& P = ctx->dp;
  const int gridL = std::min(nscans, ctx->numSms), gridG = std::min(nscans, NGLOBAL);
  const size_t smemG = cluster_smem_bytes_global(NT2);
  int* ovfR = &s.d_ctr->ovf_rings;
  int* ovfR2 = &s.d_ctr->ovf_rings2;
  int* ovfM = &s.d_ctr->ovf_merge;
  int* ovfM2 = &s.d_ctr->ovf_merge2;
  float4* kc = wantKc ? s.d_kcPool : nullptr;
  int* kcB = wantKc ? s.d_kcBase : nullptr;
  int* kcC = wantKc ? s.d_kcCnt : nullptr;
  const int sr = singleRing ? 1 : 0;
#define FE_K2_ARGS s.d_crop, s.d_cropMeta, s.d_cropCnt, s.d_scan_off, s.d_chunk_off, P, sr, s.d_kfPool, s.capKf, s.d_kfBase, s.d_kfCnt, kc, \
                   s.capKc, kcB, kcC, s.d_ctr
  // K2: the run-based kernel takes every scan; what it cannot handle (a ring with more than RW runs — unordered
  // input — or more ring entries than the scan's scratch slot) goes down the chain of grid-based instantiations.
  if (ctx->gridClustering) {
    k_cluster_rings<ECAP, NTF, 4, false><<<nscans, NTF, kClusterSmem, s.stream>>>(FE_K2_ARGS, nullptr, nullptr, s.d_ovfRings, ovfR, nullptr);
    ctx->launches++;
  } else {
#define FE_RR_ARGS s.d_crop, s.d_cropMeta, s.d_cropCnt, s.d_scan_off, s.d_chunk_off, P, sr, s.d_ringPts, s.d_ringBase, s.d_kfPool, s.capKf, \
                   s.d_kfBase, s.d_kfCnt, kc, s.capKc, kcB, kcC, s.d_ctr, s.d_ovfRuns, &s.d_ctr->ovf_runs, s.d_scanFlag, s.d_ovfRings, ovfR
    if (nscans <= GRAPH_MAX_SCANS)  // a handful of scans: a warp for each of a scan's 16 rings at once (latency, not throughput)
      k_ring_runs<NW_RR_SMALL><<<nscans, NW_RR_SMALL * 32, sizeof(RingRunsSmT<RW, NW_RR_SMALL>), s.stream>>>(FE_RR_ARGS);
    else
      k_ring_runs<NW_RR><<<nscans, NT_RR, sizeof(RingRunsSmT<RW>), s.stream>>>(FE_RR_ARGS);
#undef FE_RR_ARGS
    ctx->launches++;
    if (!lean) {
      k_ring_runs_wide<<<std::min(nscans * 4, ctx->numSms * 7), NT_RR, NW_RR * sizeof(RunBufT<RW2>), s.stream>>>(
          s.d_scan_off, P, s.d_ringPts, s.d_ringBase, s.d_kfPool, s.capKf, s.d_kfBase, s.d_kfCnt, kc, s.capKc, kcB, kcC, s.d_ctr,
          s.d_ovfRuns, &s.d_ctr->ovf_runs, s.d_scanFlag, s.d_ovfRings, ovfR);
      ctx->launches++;
    }
  }
  // lean: the fallback instantiations are left out; whatever the first kernel deferred shows in the counters and the
  // caller runs the sub-batch again with the whole chain
  if (!lean) {
    k_cluster_rings<ECAP_L, NTL, 1, false><<<gridL, NTL, kClusterSmemL2, s.stream>>>(FE_K2_ARGS, s.d_ovfRings, ovfR, s.d_ovfRings2, ovfR2, nullptr);
    k_cluster_rings<ECAP_G, NT2, 1, true><<<gridG, NT2, smemG, s.stream>>>(FE_K2_ARGS, s.d_ovfRings2, ovfR2, nullptr, nullptr, s.d_slabs);
    ctx->launches += 2;
  }
#undef FE_K2_ARGS
  if (marks) mark(ctx, s, "K2 ring clusters");
  if (merge) {
#define FE_K3_ARGS s.d_kfPool, s.d_kfBase, s.d_kfCnt, P, s.d_kpPool, (int)s.capKp, s.d_kpBase, s.d_kpCnt, s.d_ctr
    k_merge_keypoints<ECAP_M, NTM, 8, false><<<nscans, NTM, kClusterSmemM, s.stream>>>(FE_K3_ARGS, nullptr, nullptr, s.d_ovfMerge, ovfM, nullptr);
    ctx->launches++;
    if (!lean) {
      k_merge_keypoints<ECAP_L, NT2, 1, false><<<gridL, NT2, kClusterSmemL, s.stream>>>(FE_K3_ARGS, s.d_ovfMerge, ovfM, s.d_ovfMerge2, ovfM2, nullptr);
      k_merge_keypoints<ECAP_G, NT2, 1, true><<<gridG, NT2, smemG, s.stream>>>(FE_K3_ARGS, s.d_ovfMerge2, ovfM2, nullptr, nullptr, s.d_slabs);
      ctx->launches += 2;
    }
#undef FE_K3_ARGS
  }
}


// fe_enable_boundary_report: the four radius predicates of the path and the float pre-filter windows
BoundarySpec boundary_spec(const fe_ctx* ctx) {
  BoundarySpec B;
  B.eps = ctx->bndEps;
  const float r2[4] = {ctx->dp.r2f_cluster, ctx->dp.r2f_merge, ctx->dp.R2f, ctx->dp.rho2f};
  for (int k = 0; k < 4; k++) {
    B.r2f[k] = r2[k];
    B.bdist[k] = sqrt((double)r2[k]);
    B.win[k] = (float)(B.eps * (2.0 * B.bdist[k] + B.eps) * 1.01 + (double)r2[k] * 1e-6);
  }
  return B;
}

int ensure_bnd(fe_ctx* ctx, Slot& s) {
  if (!s.d_bnd) {
    CK(dalloc(&s.d_bnd, (size_t)s.capScans * 4));
    CK(halloc(&s.h_bnd, (size_t)s.capScans * 4));
  }
  return FE_OK;
}

// Enqueue the kernels of one sub-batch whose points are at d_pts.  k1flags selects what K1 does;
// `fromStage`: 0 = K1 first; 1 = crop/cropMeta/cropCnt already filled by the caller.
int enqueue_pipeline(fe_ctx* ctx, Slot& s, const float4* d_pts, int nscans, int64_t npts, int nch,
                     int k1flags, bool doDesc, bool singleRing, bool wantKc, RawLayout lay = RawLayout{nullptr, 0, 0, 0, 0},
                     const double* lateRp = nullptr, bool recordDone = true, bool lean = false) {
  DevParams& P = ctx->dp;
  s.nev = 0;
  mark(ctx, s, "begin");
  CK(cudaMemsetAsync(s.d_ctr, 0, sizeof(DevCounters), s.stream));
  if (nch > 0 && k1flags >= 0) {
    // With `lateRp` the levelling matrices (two sincos and a quaternion product per scan on the host,
    // ~0.4 ms for 10 k scans) are not staged yet: they are computed range by range — an eighth, an
    // eighth, a quarter, a half of the scans — and K1 is launched per range, so that all but the first
    // eighth of that host work runs while the GPU is already busy.  (16 equal ranges were measured:
    // slower, the per-range copy and launch calls starve the GPU.)
    const int cuts[5] = {0, nscans / 8, nscans / 4, nscans / 2, nscans};
    const int nparts = lateRp ? 4 : 1;
    for (int part = 0; part < nparts; part++) {
      const int a = lateRp ? cuts[part] : 0, b = lateRp ? cuts[part + 1] : nscans;
      if (b <= a) continue;
      if (lateRp) {
        for (int i = a; i < b; i++) leveling_matrix(lateRp[2 * i], lateRp[2 * i + 1], s.h_rot + 9 * i);
        CK(cudaMemcpyAsync(s.d_rot + 9 * (size_t)a, s.h_rot + 9 * (size_t)a, (size_t)(b - a) * 9 * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        // stage times: whatever the GPU waited for the host goes to its own line, not to K1
        if (part == 0) { s.nev = 0; mark(ctx, s, "begin"); } else mark(ctx, s, "host: levelling matrices");
      }
      const int ca = s.h_chunk_off[a], cb = s.h_chunk_off[b];
      if (cb <= ca) continue;
      if (lay.raw)
        k_level_crop_ring<true, true><<<cb - ca, 256, 0, s.stream>>>(d_pts, s.d_chunkTab, s.d_rot, P, k1flags,
                                                               s.d_surf, s.d_surfCnt, s.d_crop, s.d_cropMeta, s.d_cropCnt, nullptr, lay, ca);
      else
        k_level_crop_ring<false, true><<<cb - ca, 256, 0, s.stream>>>(d_pts, s.d_chunkTab, s.d_rot, P, k1flags,
                                                                s.d_surf, s.d_surfCnt, s.d_crop, s.d_cropMeta, s.d_cropCnt, nullptr, lay, ca);
      ctx->launches++;
      if (lateRp && part + 1 < nparts) mark(ctx, s, "K1 level+crop+ring");
    }
  }
  mark(ctx, s, "K1 level+crop+ring");
  const bool bnd = ctx->bndEps > 0.0 && nscans > 0;
  if (bnd) {
    int st = ensure_bnd(ctx, s);
    if (st) return st;
    CK(cudaMemsetAsync(s.d_bnd, 0, (size_t)nscans * 4 * sizeof(unsigned long long), s.stream));
    k_boundary_rings<<<nscans, 256, 0, s.stream>>>(s.d_crop, s.d_cropMeta, s.d_cropCnt, s.d_scan_off, s.d_chunk_off, singleRing ? 1 : 0,
                                                   boundary_spec(ctx), s.d_bnd);
    ctx->launches++;
  }
  s.sideK4a = false;
  if (doDesc) {
    int st = ensure_rowstart(ctx, s, nscans);
    if (st) return st;
  }
  launch_clustering(ctx, s, nscans, singleRing, wantKc, true, true, lean);
  mark(ctx, s, "K3 merge keypoints");
  if (bnd) {
    k_boundary_merge<<<nscans, 128, 0, s.stream>>>(s.d_kfPool, s.d_kfBase, s.d_kfCnt, P, boundary_spec(ctx), s.d_bnd);
    ctx->launches++;
  }
  if (nscans <= GRAPH_MAX_SCANS) {  // a handful of scans: offsets and ordered copy in one launch
    k_kp_offsets_gather_small<<<1, 1024, 0, s.stream>>>(s.d_kpCnt, nscans, s.d_kpOff, s.d_ctr, s.d_kpPool, s.d_kpBase, s.d_kpOut, s.d_kpScan);
    ctx->launches++;
  } else {
    k_kp_offsets<<<1, 1024, 0, s.stream>>>(s.d_kpCnt, nscans, s.d_kpOff, s.d_ctr);
    k_kp_gather<<<std::max(1, std::min(1024, (nscans * 8 + 255) / 256)), 256, 0, s.stream>>>(s.d_kpPool, s.d_kpBase, s.d_kpOff, nscans,
                                                                                              s.d_kpOut, s.d_kpScan);
    ctx->launches += 2;
  }
  mark(ctx, s, "keypoint CSR");
  if (doDesc) {
    // K4a runs after the keypoints are known: it only keeps the surface points a keypoint can reach
    launch_surface_grid(ctx, s, nscans, P, s.stream);
    mark(ctx, s, "K4a surface grid");
    const int gridKp = ctx->numSms * 8;
    launch_desc_mark(ctx, s, nscans, P, gridKp);
    mark(ctx, s, "K4b mark neighbours");
    if (bnd) {  // before K4c turns the marks in rho into densities
      k_boundary_support<<<ctx->numSms * 4, 256, 0, s.stream>>>(s.d_kpOut, s.d_kpScan, s.d_kpOff, nscans, s.d_sorted, surf_index(ctx, s),
                                                         s.d_scan_off, P, boundary_spec(ctx), s.d_bnd);
      if (npts > 0)
        k_boundary_density<<<nscans, 256, 0, s.stream>>>(s.d_sorted, surf_index(ctx, s), s.d_scan_off, P, s.d_surfN, s.d_rho,
                                                         boundary_spec(ctx), s.d_bnd);
      ctx->launches += 2;
    }
    launch_density(ctx, s, nscans, P, npts);
    mark(ctx, s, "K4c density");
    launch_desc_hist(ctx, s, nscans, P, gridKp, ctx->recordOutput, lean);
    mark(ctx, s, "K4d shape context");
  }
  if (wantKc && ctx->cloudOutputs && nscans > 0) {
    // ~cloud (src:137-139) and ~keypoint_cloud (src:133-135): per-scan counts, offsets and the ordered gather all
    // on the device; the host only learns the offsets (finalize_subbatch then copies the dense arrays out)
    const int gb = (nscans + 255) / 256;
    k_piece_counts<<<gb, 256, 0, s.stream>>>(s.d_cropCnt, s.d_chunk_off, nscans, s.d_perScan);
    k_offsets_scan<<<1, 1024, 0, s.stream>>>(s.d_perScan, nscans, s.d_outOff);
    k_gather_chunks<<<nscans, 256, 0, s.stream>>>(s.d_crop, s.d_cropCnt, s.d_scan_off, s.d_chunk_off, s.d_outOff, s.d_gather);
    k_pool16_counts<<<gb, 256, 0, s.stream>>>(s.d_kcCnt, nscans, s.d_perScan2);
    k_offsets_scan<<<1, 1024, 0, s.stream>>>(s.d_perScan2, nscans, s.d_outOff2);
    k_gather_pool16<<<nscans, 256, 0, s.stream>>>(s.d_kcPool, s.d_kcBase, s.d_kcCnt, s.d_outOff2, s.d_gather2);
    ctx->launches += 6;
    CK(cudaMemcpyAsync(s.h_cloudOff, s.d_outOff, (size_t)(nscans + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(s.h_kcOff, s.d_outOff2, (size_t)(nscans + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  }
  CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(s.h_kpOff, s.d_kpOff, (size_t)(nscans + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  if (bnd) CK(cudaMemcpyAsync(s.h_bnd, s.d_bnd, (size_t)nscans * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
  if (recordDone) CK(cudaEventRecord(s.evDone, s.stream));
  CK(cudaGetLastError());
  return FE_OK;
}

void collect_times(fe_ctx* ctx, Slot& s) {
  ctx->stName.clear();
  ctx->stMs.clear();
  for (int i = 1; i < s.nev; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.ev[i - 1], s.ev[i]) == cudaSuccess) {
      ctx->stName.push_back(s.evName[i]);
      ctx->stMs.push_back(ms);
    }
  }
}

int grow_results(fe_ctx* ctx, int64_t needKp, bool desc) {
  if (needKp > ctx->capResKp) {
    for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
    const int64_t cap = std::max<int64_t>(needKp * 3 / 2, 4096);
    fe_point_t* nk = nullptr;
    CK(halloc(&nk, (size_t)cap));
    if (ctx->h_kp) { memcpy(nk, ctx->h_kp, (size_t)ctx->capResKp * sizeof(fe_point_t)); cudaFreeHost(ctx->h_kp); }
    ctx->h_kp = nk;
    ctx->capResKp = cap;
  }
  if (desc && ctx->capResKp > ctx->capResDesc) {
    for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
    float* nd = nullptr;
    CK(halloc(&nd, (size_t)ctx->capResKp * FE_RECORD_FLOATS));
    if (ctx->h_desc) { memcpy(nd, ctx->h_desc, (size_t)ctx->capResDesc * FE_RECORD_FLOATS * sizeof(float)); cudaFreeHost(ctx->h_desc); }
    ctx->h_desc = nd;
    ctx->capResDesc = ctx->capResKp;
  }
  return FE_OK;
}

int grow_pinned_points(fe_ctx* ctx, fe_point_t** buf, int64_t* cap, int64_t used, int64_t need) {
  if (need <= *cap) return FE_OK;
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  const int64_t ncap = std::max<int64_t>(need * 3 / 2, 1 << 16);
  fe_point_t* nb = nullptr;
  CK(halloc(&nb, (size_t)ncap));
  if (*buf) { memcpy(nb, *buf, (size_t)used * sizeof(fe_point_t)); cudaFreeHost(*buf); }
  *buf = nb;
  *cap = ncap;
  return FE_OK;
}

// wait for the sub-batch in `s`, check its error word, append its keypoints to the results
// The sub-batch ran the lean chain (first instantiation of every stage only, see enqueue_small).  Anything one of
// them deferred to a fallback kernel is in the counters: run the sub-batch again, eagerly, with the whole chain (its
// points and staged offsets are still where they were), and keep the lean chain off for a while — inputs that defer
// once (unordered clouds, very large descriptor radii) usually keep doing so.  Call after the slot's evDone.
int lean_rerun_if_deferred(fe_ctx* ctx, Slot& s) {
  if (!s.lean) return FE_OK;
  s.lean = false;
  const DevCounters& c = *s.h_ctr;
  if (c.err || !(c.ovf_runs | c.ovf_rings | c.ovf_rings2 | c.ovf_merge | c.ovf_merge2 | c.n_list_m | c.n_list_l)) return FE_OK;
  ctx->leanHold = 256;
  ctx->leanReruns++;
  RawLayout lay = {s.rrRaw ? (const unsigned char*)s.rrPts : nullptr, s.rrStride, s.rrXo, s.rrYo, s.rrZo};
  int st = enqueue_pipeline(ctx, s, s.rrPts, s.nscans, s.npts, s.rrNch, s.rrK1flags, s.rrDesc, false, s.rrWantKc, lay);
  if (st) return st;
  CK(cudaEventSynchronize(s.evDone));
  return FE_OK;
}

// A sub-batch of at most GRAPH_MAX_SCANS scans whose offsets stage_scans(copy = false) left in the pinned mirrors:
// the chain is captured into a CUDA graph per shape (second sighting) and replayed from then on.  Lean chain: a
// handful of scans almost never needs a fallback instantiation, and seven kernels that only find that out cost more
// than a tenth of a single scan's latency; they are left out and lean_rerun_if_deferred repairs the rare miss.
// Records the slot's evDone.
int enqueue_small(fe_ctx* ctx, Slot& s, const float4* d_pts, bool ptsInKey, int ns, int64_t npts, int nch, int k1flags, bool desc,
                  bool wantKc, RawLayout lay) {
  int st;
  // every allocation the pipeline may need happens before a capture starts
  if (desc) { st = ensure_rowstart(ctx, s, ns); if (st) return st; }
  if (ctx->bndEps > 0.0) { st = ensure_bnd(ctx, s); if (st) return st; }
  GraphKey key;
  memset(&key, 0, sizeof key);
  key.nscans = ns; key.nch = nch; key.by = density_blocks_per_scan(ns, npts); key.k1flags = k1flags | ((s.maxScanPts > 65535) ? (1 << 16) : 0);
  key.stride = lay.raw ? lay.stride : 0; key.xo = lay.xo; key.yo = lay.yo; key.zo = lay.zo;
  key.desc = desc; key.wantKc = wantKc; key.bnd = ctx->bndEps > 0.0; key.epoch = ctx->epoch;
  key.pts = ptsInKey ? (const void*)d_pts : nullptr;
  const bool lean = ctx->leanHold == 0;
  if (ctx->leanHold > 0) ctx->leanHold--;
  key.lean = lean;
  s.lean = lean; s.rrPts = d_pts; s.rrNch = nch; s.rrK1flags = k1flags; s.rrDesc = desc; s.rrWantKc = wantKc;
  s.rrRaw = lay.raw != nullptr; s.rrStride = lay.stride; s.rrXo = lay.xo; s.rrYo = lay.yo; s.rrZo = lay.zo;
  GraphEntry* ge = nullptr;
  for (GraphEntry& g : s.graphs) if (g.key == key) { ge = &g; break; }
  if (ge && ge->exec) {  // replay
    CK(cudaGraphLaunch(ge->exec, s.stream));
    ctx->launches += ge->launches;
    ctx->graphReplays++;
  } else if (!ge) {      // first sighting of the shape: eager, so that every lazily initialised piece exists
    st = stage_scans_copy(ctx, s, ns, false);
    if (st) return st;
    st = enqueue_pipeline(ctx, s, d_pts, ns, npts, nch, k1flags, desc, false, wantKc, lay, nullptr, false, lean);
    if (st) return st;
    if (s.graphs.size() >= 16) {  // drop the oldest shape
      if (s.graphs[0].exec) cudaGraphExecDestroy(s.graphs[0].exec);
      s.graphs.erase(s.graphs.begin());
    }
    GraphEntry e; e.key = key;
    s.graphs.push_back(e);
  } else {               // second sighting: capture, instantiate, launch
    const int64_t l0 = ctx->launches;
    CK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
    st = stage_scans_copy(ctx, s, ns, false);
    if (st == FE_OK) st = enqueue_pipeline(ctx, s, d_pts, ns, npts, nch, k1flags, desc, false, wantKc, lay, nullptr, false, lean);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s.stream, &graph);
    if (st) { if (graph) cudaGraphDestroy(graph); return st; }
    if (ce != cudaSuccess) return fail(ctx, FE_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess) return fail(ctx, FE_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ci));
    ge->exec = exec;
    ge->launches = ctx->launches - l0;
    CK(cudaGraphLaunch(exec, s.stream));
    ctx->graphReplays++;
  }
  CK(cudaEventRecord(s.evDone, s.stream));
  return FE_OK;
}

int finalize_subbatch(fe_ctx* ctx, Slot& s, int64_t& kpRun, bool desc) {
  if (!s.busy) return FE_OK;
  CK(cudaEventSynchronize(s.evDone));
  s.busy = false;
  { int st = lean_rerun_if_deferred(ctx, s); if (st) return st; }
  if (s.h_ctr->err) return fail(ctx, FE_ERR_CAPACITY, err_bits(s.h_ctr->err));
  const int K = s.h_kpOff[s.nscans];
  int st = grow_results(ctx, kpRun + K, desc);
  if (st) return st;
  for (int i = 0; i <= s.nscans; i++) ctx->kpOffsets[s.firstScan + i] = kpRun + s.h_kpOff[i];
  if (ctx->bndEps > 0.0 && s.h_bnd)
    for (int i = 0; i < s.nscans * 4; i++) ctx->bndCounts[(size_t)s.firstScan * 4 + i] = (int64_t)s.h_bnd[i];
  if (K > 0) {
    CK(cudaMemcpyAsync(ctx->h_kp + kpRun, s.d_kpOut, (size_t)K * sizeof(float4), cudaMemcpyDeviceToHost, s.stream));
    const size_t dl = ctx->recordOutput ? FE_RECORD_FLOATS : FE_DESC_LEN;
    if (desc) CK(cudaMemcpyAsync(ctx->h_desc + kpRun * dl, s.d_desc, (size_t)K * dl * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
  }
  kpRun += K;
  if (ctx->cloudOutputs && s.h_cloudOff) {  // the two optional clouds, gathered on the device by enqueue_pipeline
    const int64_t tc = s.h_cloudOff[s.nscans], tk = s.h_kcOff[s.nscans];
    st = grow_pinned_points(ctx, &ctx->h_cloud, &ctx->capCloud, ctx->cloudRun, ctx->cloudRun + tc);
    if (st) return st;
    st = grow_pinned_points(ctx, &ctx->h_kc, &ctx->capKcRes, ctx->kcRun, ctx->kcRun + tk);
    if (st) return st;
    for (int i = 0; i <= s.nscans; i++) {
      ctx->cloudOff[s.firstScan + i] = ctx->cloudRun + s.h_cloudOff[i];
      ctx->kcOff[s.firstScan + i] = ctx->kcRun + s.h_kcOff[i];
    }
    if (tc > 0) CK(cudaMemcpyAsync(ctx->h_cloud + ctx->cloudRun, s.d_gather, (size_t)tc * sizeof(float4), cudaMemcpyDeviceToHost, s.stream));
    if (tk > 0) CK(cudaMemcpyAsync(ctx->h_kc + ctx->kcRun, s.d_gather2, (size_t)tk * sizeof(float4), cudaMemcpyDeviceToHost, s.stream));
    ctx->cloudRun += tc;
    ctx->kcRun += tk;
  }
  return FE_OK;
}

int gather_to_host(fe_ctx* ctx, Slot& s, int nscans, bool chunks, const float4* src, const int* cnt, const int* pBase,
                   std::vector<int64_t>& offOut, std::vector<fe_point_t>& ptsOut) {
  if (chunks) k_piece_counts<<<(nscans + 255) / 256, 256, 0, s.stream>>>(cnt, s.d_chunk_off, nscans, s.d_perScan);
  else k_pool16_counts<<<(nscans + 255) / 256, 256, 0, s.stream>>>(cnt, nscans, s.d_perScan);
  ctx->launches++;
  CK(cudaMemcpyAsync(s.h_perScan, s.d_perScan, (size_t)nscans * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  offOut.assign(nscans + 1, 0);
  std::vector<int> off32(nscans + 1, 0);
  for (int i = 0; i < nscans; i++) { offOut[i + 1] = offOut[i] + s.h_perScan[i]; off32[i + 1] = (int)offOut[i + 1]; }
  const int64_t tot = offOut[nscans];
  ptsOut.resize((size_t)tot);
  if (tot == 0) return FE_OK;
  if (tot > s.capPts) return fail(ctx, FE_ERR_CAPACITY, "cloud output exceeds max_points_per_call");
  CK(cudaMemcpyAsync(s.d_outOff, off32.data(), (size_t)(nscans + 1) * sizeof(int), cudaMemcpyHostToDevice, s.stream));
  if (chunks) k_gather_chunks<<<nscans, 256, 0, s.stream>>>(src, cnt, s.d_scan_off, s.d_chunk_off, s.d_outOff, s.d_gather);
  else k_gather_pool16<<<nscans, 256, 0, s.stream>>>(src, pBase, cnt, s.d_outOff, s.d_gather);
  ctx->launches++;
  CK(cudaMemcpyAsync(ptsOut.data(), s.d_gather, (size_t)tot * sizeof(float4), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return FE_OK;
}

// A host batch call that fails half-way leaves sub-batches in flight on the two slots.  Whatever the
// exit path, both streams are drained and the slots marked idle, so the next call on the context starts
// clean instead of finalising a stale sub-batch into its own result arrays.
struct SlotDrain {
  fe_ctx* ctx;
  bool ok = false;
  explicit SlotDrain(fe_ctx* c) : ctx(c) { drain(); }
  ~SlotDrain() { if (!ok) drain(); }
  void drain() {
    for (int k = 0; k < 2; k++) {
      Slot& s = ctx->slot[k];
      if (s.stream) cudaStreamSynchronize(s.stream);
      if (s.stream2) cudaStreamSynchronize(s.stream2);
      s.busy = false;
    }
  }
};

}  // namespace

// ================================================================================================
extern "C" {

const char* fe_version(void) { return FE_VERSION; }

void fe_params_node_default(fe_params_t* p) {
  p->x_min = 0.0; p->x_max = 75.0;
  p->y_min = -30.0; p->y_max = 30.0;
  p->z_min = -1.5; p->z_max = 5.0;
  p->cluster_tolerance = 0.65;
  p->cluster_min_count = 5;
  p->cluster_max_count = 50;
  p->cluster_radius_threshold = 0.15;
  p->number_detection_channels = 1;
  p->estimate_descriptors = 1;
  p->descriptor_radius = 2.5;
}

void fe_params_launch_playback(fe_params_t* p) {
  fe_params_node_default(p);
  p->x_max = 100.0;
  p->y_min = -50.0; p->y_max = 50.0;
  p->z_max = 4.0;
  p->cluster_tolerance = 1.0;
  p->cluster_min_count = 1;
  p->cluster_max_count = 1000;
  p->cluster_radius_threshold = 0.2;
  p->number_detection_channels = 2;
}

int fe_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void* fe_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void fe_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* fe_last_error(const fe_ctx_t* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fe_create(int device, const fe_params_t* params, const fe_limits_t* limits, fe_ctx_t** out) {
  if (!out || !params) return FE_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return FE_ERR_NO_DEVICE;  // no CPU fallback
  if (device < 0 || device >= n) return FE_ERR_INVALID;
  fe_ctx* ctx = new fe_ctx();
  ctx->device = device;
  fe_limits_t L = {0, 0, 0, 0};
  if (limits) L = *limits;
  if (L.max_points_per_call <= 0) L.max_points_per_call = 32LL << 20;
  if (L.max_scans_per_call <= 0) L.max_scans_per_call = 2048;
  if (L.max_keypoints_per_call <= 0) L.max_keypoints_per_call = 64 << 10;
  if (L.max_ring_clusters_per_call <= 0) L.max_ring_clusters_per_call = 1 << 20;
  if (L.max_points_per_call >= (1LL << 32) - CH) { delete ctx; return FE_ERR_INVALID; }
  ctx->lim = L;
  auto bail = [&](int code) { fe_destroy(ctx); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(FE_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(FE_ERR_CUDA);
  if (prop.major != 10) return bail(FE_ERR_NO_DEVICE);  // built for sm_100a only
  ctx->numSms = std::max(prop.multiProcessorCount, 1);
  if (cudaMalloc((void**)&ctx->d_lut, FE_DESC_LEN * sizeof(float)) != cudaSuccess) return bail(FE_ERR_CUDA);
  // 3DSC x-axes: boost::uniform_01<mt19937>(12345) draws 3k..3k+2 -> (x0, x1, -0) normalised
  ctx->axesCap = 1 << 16;
  {
    std::vector<float2> ax(ctx->axesCap);
    std::mt19937 gen(12345u);
    for (int k = 0; k < ctx->axesCap; k++) {
      const float u0 = (float)((double)gen() * (1.0 / 4294967296.0));
      const float u1 = (float)((double)gen() * (1.0 / 4294967296.0));
      (void)gen();  // the third draw is overwritten by -(n.x*x0 + n.y*x1)/n.z = -0
      const float z = -(0.0f * u0 + 0.0f * u1) / 1.0f;
      const float inv = 1.0f / sqrtf(u0 * u0 + (u1 * u1 + z * z));
      ax[k].x = u0 * inv;
      ax[k].y = u1 * inv;
    }
    if (cudaMalloc((void**)&ctx->d_axes, ax.size() * sizeof(float2)) != cudaSuccess) return bail(FE_ERR_CUDA);
    if (cudaMemcpy(ctx->d_axes, ax.data(), ax.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) return bail(FE_ERR_CUDA);
  }
  int st = derive_params(ctx, *params);
  if (st) { fe_destroy(ctx); return st; }
  st = set_kernel_attrs(ctx);
  if (st) { fe_destroy(ctx); return st; }
  *out = ctx;
  return FE_OK;
}

int fe_set_params(fe_ctx_t* ctx, const fe_params_t* params) {
  if (!ctx || !params) return FE_ERR_INVALID;
  cudaSetDevice(ctx->device);
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) cudaStreamSynchronize(ctx->slot[k].stream);
  ctx->epoch++;
  return derive_params(ctx, *params);
}

void fe_destroy(fe_ctx_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int k = 0; k < 2; k++) free_slot(ctx->slot[k]);
  if (ctx->d_lut) cudaFree(ctx->d_lut);
  if (ctx->d_axes) cudaFree(ctx->d_axes);
  if (ctx->h_kp) cudaFreeHost(ctx->h_kp);
  if (ctx->h_desc) cudaFreeHost(ctx->h_desc);
  if (ctx->h_cloud) cudaFreeHost(ctx->h_cloud);
  if (ctx->h_kc) cudaFreeHost(ctx->h_kc);
  delete ctx;
}

int fe_enable_cloud_outputs(fe_ctx_t* ctx, int32_t enable) {
  if (!ctx) return FE_ERR_INVALID;
  ctx->cloudOutputs = enable != 0;
  ctx->epoch++;
  return FE_OK;
}

int fe_enable_stage_timing(fe_ctx_t* ctx, int32_t enable) {
  if (!ctx) return FE_ERR_INVALID;
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  ctx->stageTiming = enable != 0;
  ctx->epoch++;
  return FE_OK;
}

int fe_enable_record_output(fe_ctx_t* ctx, int32_t enable) {
  if (!ctx) return FE_ERR_INVALID;
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  ctx->recordOutput = enable != 0;
  ctx->epoch++;
  return FE_OK;
}

int fe_set_angle_libm(fe_ctx_t* ctx, int32_t mode) {
  if (!ctx || (mode != FE_LIBM_FDLIBM && mode != FE_LIBM_CORRECTLY_ROUNDED)) return FE_ERR_INVALID;
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  ctx->angleLibm = mode;
  ctx->dp.angle_libm = mode;
  ctx->epoch++;
  return FE_OK;
}

int fe_enable_boundary_report(fe_ctx_t* ctx, double eps_m) {
  if (!ctx || !(eps_m == eps_m)) return FE_ERR_INVALID;
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  ctx->bndEps = eps_m > 0.0 ? eps_m : 0.0;
  ctx->bndCounts.clear();
  ctx->epoch++;
  return FE_OK;
}

int fe_get_boundary_report(fe_ctx_t* ctx, const int64_t** counts, int32_t* n_scans) {
  if (!ctx || !counts || !n_scans) return FE_ERR_INVALID;
  if (!(ctx->bndEps > 0.0)) return fail(ctx, FE_ERR_INVALID, "boundary report not enabled (fe_enable_boundary_report)");
  *counts = ctx->bndCounts.data();
  *n_scans = (int32_t)(ctx->bndCounts.size() / 4);
  return FE_OK;
}

int fe_debug_libm_f32(fe_ctx_t* ctx, int32_t op, const float* a, const float* b, float* out, int64_t n) {
  if (!ctx || op < 0 || op > 2 || n < 0 || (n > 0 && (!a || !out || (op == 0 && !b)))) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return FE_OK;
  float *da = nullptr, *db = nullptr, *dout = nullptr;
  auto done = [&](int code) { cudaFree(da); cudaFree(db); cudaFree(dout); return code; };
  if (cudaMalloc((void**)&da, (size_t)n * 4) != cudaSuccess || cudaMalloc((void**)&db, (size_t)n * 4) != cudaSuccess ||
      cudaMalloc((void**)&dout, (size_t)n * 4) != cudaSuccess)
    return done(fail(ctx, FE_ERR_CUDA, "fe_debug_libm_f32: out of device memory"));
  cudaMemcpy(da, a, (size_t)n * 4, cudaMemcpyHostToDevice);
  if (op == 0) cudaMemcpy(db, b, (size_t)n * 4, cudaMemcpyHostToDevice);
  k_debug_libm_f32<<<ctx->numSms * 8, 256>>>(op, da, db, dout, (long long)n);
  ctx->launches++;
  if (cudaMemcpy(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return done(fail(ctx, FE_ERR_CUDA, std::string("fe_debug_libm_f32: ") + cudaGetErrorString(cudaGetLastError())));
  return done(FE_OK);
}

int fe_get_cloud_outputs(fe_ctx_t* ctx, const int64_t** cloud_offsets, const fe_point_t** cloud,
                         const int64_t** kpcloud_offsets, const fe_point_t** keypoint_cloud) {
  if (!ctx) return FE_ERR_INVALID;
  if (ctx->cloudOff.empty()) return fail(ctx, FE_ERR_INVALID, "no cloud outputs recorded (fe_enable_cloud_outputs before fe_process_batch)");
  if (cloud_offsets) *cloud_offsets = ctx->cloudOff.data();
  if (cloud) *cloud = ctx->h_cloud;
  if (kpcloud_offsets) *kpcloud_offsets = ctx->kcOff.data();
  if (keypoint_cloud) *keypoint_cloud = ctx->h_kc;
  return FE_OK;
}

int fe_get_stage_times(fe_ctx_t* ctx, int32_t cap, const char** names, float* ms, int32_t* n) {
  if (!ctx || !n || cap < 0 || (cap > 0 && (!names || !ms))) return FE_ERR_INVALID;
  const int m = std::min<int>(cap, (int)ctx->stName.size());
  for (int i = 0; i < m; i++) { names[i] = ctx->stName[i]; ms[i] = ctx->stMs[i]; }
  *n = m;
  return FE_OK;
}

// ---- the fused path ------------------------------------------------------------------------------

static int process_batch_host(fe_ctx_t* ctx, const unsigned char* points, int stride, int xo, int yo, int zo, bool isFloat4,
                              const int64_t* scan_offsets, const double* roll_pitch, int32_t n_scans, fe_batch_result_t* out) {
  if (!ctx || !out || n_scans < 0 || (n_scans > 0 && (!scan_offsets || !roll_pitch))) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  ctx->err.clear();
  if (n_scans > 0 && scan_offsets[0] < 0) return fail(ctx, FE_ERR_INVALID, "scan_offsets must not be negative");
  SlotDrain guard(ctx);
  const bool desc = ctx->params.estimate_descriptors != 0;
  const int64_t launches0 = ctx->launches;
  ctx->kpOffsets.assign((size_t)n_scans + 1, 0);
  ctx->bndCounts.assign(ctx->bndEps > 0.0 ? (size_t)n_scans * 4 : 0, 0);
  ctx->cloudOff.clear(); ctx->kcOff.clear();
  ctx->cloudRun = ctx->kcRun = 0;
  if (ctx->cloudOutputs) { ctx->cloudOff.assign((size_t)n_scans + 1, 0); ctx->kcOff.assign((size_t)n_scans + 1, 0); }
  int64_t kpRun = 0;
  int first = 0, cur = 0;
  int nsub = 0;
  while (first < n_scans) {
    Slot& s = ctx->slot[cur];
    int st = ensure_slot(ctx, s, true);
    if (st) return st;
    // previous sub-batch of this slot must have been finalised (its device buffers are reused)
    st = finalize_subbatch(ctx, s, kpRun, desc);
    if (st) return st;
    // greedy sub-batch
    int last = first;
    int64_t npts = 0;
    while (last < n_scans && (last - first) < s.capScans) {
      const int64_t n = scan_offsets[last + 1] - scan_offsets[last];
      if (n < 0) return fail(ctx, FE_ERR_INVALID, "scan_offsets must be non-decreasing");
      if (n > s.capPts || n * stride > s.capPts * 16) return fail(ctx, FE_ERR_CAPACITY, "a single scan exceeds max_points_per_call");
      if (npts + n > s.capPts || (npts + n) * stride > s.capPts * 16) break;
      npts += n;
      last++;
    }
    const int ns = last - first;
    int64_t np2; int nch;
    const bool graphable = ctx->useGraphs && !ctx->stageTiming && ns <= GRAPH_MAX_SCANS;
    st = stage_scans(ctx, s, scan_offsets + first, roll_pitch + 2 * first, ns, &np2, &nch, false, !graphable);
    if (st) return st;
    if (npts > 0)
      CK(cudaMemcpyAsync(s.d_pts, points + scan_offsets[first] * (int64_t)stride, (size_t)npts * (size_t)stride, cudaMemcpyHostToDevice, s.stream));
    const bool wantKc = ctx->cloudOutputs;
    if (wantKc) { st = ensure_kc(ctx, s); if (st) return st; }
    RawLayout lay = {isFloat4 ? nullptr : (const unsigned char*)s.d_pts, stride, xo, yo, zo};
    const int k1flags = F_ELEV | F_ROT | F_CROP | F_RING | (desc ? F_SURF : 0);
    if (!graphable) {
      s.lean = false;
      st = enqueue_pipeline(ctx, s, s.d_pts, ns, npts, nch, k1flags, desc, false, wantKc, lay);
      if (st) return st;
    } else {
      st = enqueue_small(ctx, s, s.d_pts, false, ns, npts, nch, k1flags, desc, wantKc, lay);
      if (st) return st;
    }
    s.busy = true; s.nscans = ns; s.npts = npts; s.firstScan = first;
    nsub++;
    // the other slot's sub-batch was enqueued earlier: finalise it while this one runs
    Slot& o = ctx->slot[cur ^ 1];
    if (o.busy) { st = finalize_subbatch(ctx, o, kpRun, desc); if (st) return st; }
    first = last;
    cur ^= 1;
  }
  // drain in submission order
  for (int k = 0; k < 2; k++) {
    Slot& s = ctx->slot[cur ^ k];  // cur now points at the older one
    int st = finalize_subbatch(ctx, s, kpRun, desc);
    if (st) return st;
  }
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  if (nsub == 1) collect_times(ctx, ctx->slot[0]);
  out->n_scans = n_scans;
  out->n_keypoints = kpRun;
  out->keypoint_offsets = ctx->kpOffsets.data();
  out->keypoints = ctx->h_kp;
  out->descriptors = desc ? ctx->h_desc : nullptr;
  out->on_device = 0;
  out->gpu_launches = ctx->launches - launches0;
  guard.ok = true;
  return FE_OK;
}

int fe_process_batch(fe_ctx_t* ctx, const fe_point_t* points, const int64_t* scan_offsets,
                     const double* roll_pitch, int32_t n_scans, fe_batch_result_t* out) {
  return process_batch_host(ctx, (const unsigned char*)points, 16, 0, 4, 8, true, scan_offsets, roll_pitch, n_scans, out);
}

int fe_process_batch_layout(fe_ctx_t* ctx, const void* points, const fe_point_layout_t* layout,
                            const int64_t* scan_offsets, const double* roll_pitch, int32_t n_scans,
                            fe_batch_result_t* out) {
  if (!ctx || !layout) return FE_ERR_INVALID;
  const fe_point_layout_t& L = *layout;
  if (L.stride < 12 || L.stride > 4096 || L.x_off < 0 || L.y_off < 0 || L.z_off < 0 || L.x_off + 4 > L.stride ||
      L.y_off + 4 > L.stride || L.z_off + 4 > L.stride)
    return fail(ctx, FE_ERR_INVALID, "fe_point_layout: need 12 <= stride <= 4096 and x/y/z floats inside the record");
  return process_batch_host(ctx, (const unsigned char*)points, L.stride, L.x_off, L.y_off, L.z_off, false, scan_offsets,
                            roll_pitch, n_scans, out);
}

// tf::Matrix3x3(quat).getRPY(tmproll, pitch, yaw) as imuCallback uses it (src:60-65)
int fe_imu_to_roll_pitch(const double q[4], int32_t cloud_leveling, double* roll, double* pitch) {
  if (!q || !roll || !pitch) return FE_ERR_INVALID;
  if (!cloud_leveling) { *roll = 0.0; *pitch = 0.0; return FE_OK; }  // src:66-69
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  // Matrix3x3::setRotation
  const double d = x * x + y * y + z * z + w * w;
  const double s = 2.0 / d;
  const double xs = x * s, ys = y * s;
  const double wx = w * xs, wy = w * ys;
  const double xx = x * xs, xz = x * (z * s);
  const double yy = y * ys, yz = y * (z * s);
  const double m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  // Matrix3x3::getEulerYPR, solution 1
  double tmproll, p;
  if (fabs(m20) >= 1.0) {
    const double delta = atan2(m21, m22);
    p = (m20 < 0.0) ? 3.14159265358979323846 / 2.0 : -3.14159265358979323846 / 2.0;
    tmproll = delta;
  } else {
    p = -asin(m20);
    tmproll = atan2(m21 / cos(p), m22 / cos(p));
  }
  *roll = tmproll - 3.14159265358979323846;  // src:65
  *pitch = p;
  return FE_OK;
}

int fe_process_batch_device(fe_ctx_t* ctx, const fe_point_t* d_points, const int64_t* scan_offsets,
                            const double* roll_pitch, int32_t n_scans, fe_batch_result_t* out) {
  if (!ctx || !out || n_scans < 0 || (n_scans > 0 && (!scan_offsets || !roll_pitch || !d_points))) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  ctx->err.clear();
  if (n_scans > 0 && scan_offsets[0] < 0) return fail(ctx, FE_ERR_INVALID, "scan_offsets must not be negative");
  const bool desc = ctx->params.estimate_descriptors != 0;
  const int64_t launches0 = ctx->launches;
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, false);
  if (st) return st;
  ctx->kpOffsets.assign((size_t)n_scans + 1, 0);
  int64_t npts = 0; int nch = 0;
  if (n_scans > 0) {
    const bool late = n_scans >= 2048;
    const bool small = ctx->useGraphs && !ctx->stageTiming && n_scans <= GRAPH_MAX_SCANS;
    const int k1flags = F_ELEV | F_ROT | F_CROP | F_RING | (desc ? F_SURF : 0);
    const float4* d_pts = (const float4*)d_points + scan_offsets[0];
    st = stage_scans(ctx, s, scan_offsets, roll_pitch, n_scans, &npts, &nch, late, !small);
    if (st) return st;
    s.nscans = n_scans; s.npts = npts;
    if (small) {  // a handful of scans: lean chain, replayed from a graph captured for this shape and input pointer
      st = enqueue_small(ctx, s, d_pts, true, n_scans, npts, nch, k1flags, desc, false, RawLayout{nullptr, 0, 0, 0, 0});
    } else {
      s.lean = false;
      st = enqueue_pipeline(ctx, s, d_pts, n_scans, npts, nch, k1flags, desc, false, false, RawLayout{nullptr, 0, 0, 0, 0},
                            late ? roll_pitch : nullptr);
    }
    if (st) return st;
    CK(cudaEventSynchronize(s.evDone));
    st = lean_rerun_if_deferred(ctx, s);
    if (st) return st;
    if (s.h_ctr->err) return fail(ctx, FE_ERR_CAPACITY, err_bits(s.h_ctr->err));
    for (int i = 0; i <= n_scans; i++) ctx->kpOffsets[i] = s.h_kpOff[i];
    ctx->bndCounts.assign(ctx->bndEps > 0.0 ? (size_t)n_scans * 4 : 0, 0);
    if (ctx->bndEps > 0.0 && s.h_bnd)
      for (int i = 0; i < n_scans * 4; i++) ctx->bndCounts[i] = (int64_t)s.h_bnd[i];
    collect_times(ctx, s);
    s.nscans = n_scans; s.npts = npts; s.lastNch = nch;
  }
  out->n_scans = n_scans;
  out->n_keypoints = ctx->kpOffsets[n_scans];
  out->keypoint_offsets = ctx->kpOffsets.data();
  out->keypoints = (const fe_point_t*)s.d_kpOut;
  out->descriptors = desc ? s.d_desc : nullptr;
  out->on_device = 1;
  out->gpu_launches = ctx->launches - launches0;
  return FE_OK;
}

int fe_timer_begin(fe_ctx_t* ctx) {
  if (!ctx) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int st = ensure_slot(ctx, ctx->slot[0], false);
  if (st) return st;
  CK(cudaEventRecord(ctx->slot[0].evT0, ctx->slot[0].stream));
  return FE_OK;
}

int fe_timer_end(fe_ctx_t* ctx, float* elapsed_ms) {
  if (!ctx || !elapsed_ms || !ctx->slot[0].stream) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->slot[0].evT1, ctx->slot[0].stream));
  CK(cudaEventSynchronize(ctx->slot[0].evT1));
  CK(cudaEventElapsedTime(elapsed_ms, ctx->slot[0].evT0, ctx->slot[0].evT1));
  return FE_OK;
}

int fe_get_batch_stats(fe_ctx_t* ctx, int64_t out[10]) {
  if (!ctx || !out || !ctx->slot[0].stream) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  CK(cudaStreamSynchronize(s.stream));
  for (int i = 0; i < 10; i++) out[i] = 0;
  out[0] = s.npts;
  std::vector<int> a((size_t)std::max(s.lastNch, 1)), b((size_t)std::max(s.lastNch, 1));
  if (s.lastNch > 0) {
    CK(cudaMemcpy(a.data(), s.d_surfCnt, (size_t)s.lastNch * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), s.d_cropCnt, (size_t)s.lastNch * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < s.lastNch; i++) { out[1] += a[i]; out[2] += b[i]; }
  }
  if (s.nscans > 0) {
    std::vector<int> kf((size_t)s.nscans * 16);
    CK(cudaMemcpy(kf.data(), s.d_kfCnt, kf.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int v : kf) out[3] += v;
    const int K = s.h_kpOff[s.nscans];
    out[4] = K;
    if (K > 0 && ctx->params.estimate_descriptors) {
      std::vector<int> nb((size_t)K);
      CK(cudaMemcpy(nb.data(), s.d_kpNbr, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost));
      for (int v : nb) out[5] += v;
    }
  }
  out[6] = s.h_ctr->ovf_rings;  // (of which ovf_rings2 went on to the large instantiation)
  out[7] = s.h_ctr->ovf_merge;
  out[8] = s.h_ctr->ovf_surf;
  out[9] = s.h_ctr->desc_unordered;
  return FE_OK;
}

int fe_download(fe_ctx_t* ctx, void* host_dst, const void* device_src, int64_t bytes) {
  if (!ctx || bytes < 0 || (bytes > 0 && (!host_dst || !device_src))) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (bytes > 0) CK(cudaMemcpy(host_dst, device_src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return FE_OK;
}

// ---- one entry point per reference function ----------------------------------------------------------

static int stage_k1(fe_ctx* ctx, fe_point_t* cloud, int64_t n, double roll, double pitch, int flags) {
  // K1 with full_out: every point transformed in place (no compaction)
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, true);
  if (st) return st;
  if (n == 0) return FE_OK;
  if (!s.d_full) CK(dalloc(&s.d_full, (size_t)s.capPts));
  int64_t offs[2] = {0, n};
  double rp[2] = {roll, pitch};
  int64_t npts; int nch;
  st = stage_scans(ctx, s, offs, rp, 1, &npts, &nch);
  if (st) return st;
  CK(cudaMemcpyAsync(s.d_pts, cloud, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, s.stream));
  k_level_crop_ring<false, false><<<nch, 256, 0, s.stream>>>(s.d_pts, s.d_chunkTab, s.d_rot, ctx->dp, flags, s.d_surf,
                                               s.d_surfCnt, s.d_crop, s.d_cropMeta, s.d_cropCnt, s.d_full, RawLayout{nullptr, 0, 0, 0, 0}, 0);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(cloud, s.d_full, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return FE_OK;
}

int fe_get_elevation_angles(fe_ctx_t* ctx, fe_point_t* cloud, int64_t n) {
  if (!ctx || (n > 0 && !cloud) || n < 0) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return stage_k1(ctx, cloud, n, 0.0, 0.0, F_ELEV);
}

int fe_rotate_cloud(fe_ctx_t* ctx, fe_point_t* cloud, int64_t n, double roll, double pitch) {
  if (!ctx || (n > 0 && !cloud) || n < 0) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return stage_k1(ctx, cloud, n, roll, pitch, F_ROT);
}

int fe_rotation_matrix(double roll, double pitch, float m[9]) {
  if (!m) return FE_ERR_INVALID;
  leveling_matrix(roll, pitch, m);
  return FE_OK;
}

// upload a one-scan cloud and run K1 with the given flags (compacting into the chunk pieces)
static int stage_upload_k1(fe_ctx* ctx, Slot& s, const fe_point_t* in, int64_t n, int flags, int* nchOut) {
  int64_t offs[2] = {0, n};
  int64_t npts;
  int st = stage_scans(ctx, s, offs, nullptr, 1, &npts, nchOut);
  if (st) return st;
  if (n > 0) CK(cudaMemcpyAsync(s.d_pts, in, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, s.stream));
  if (*nchOut > 0) {
    k_level_crop_ring<false, false><<<*nchOut, 256, 0, s.stream>>>(s.d_pts, s.d_chunkTab, s.d_rot, ctx->dp, flags, s.d_surf,
                                                     s.d_surfCnt, s.d_crop, s.d_cropMeta, s.d_cropCnt, nullptr, RawLayout{nullptr, 0, 0, 0, 0}, 0);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  return FE_OK;
}

static int copy_points_out(fe_ctx* ctx, const std::vector<fe_point_t>& v, fe_point_t* out, int64_t cap, int64_t* n_out) {
  if (n_out) *n_out = (int64_t)v.size();
  if (!out) return FE_OK;
  if ((int64_t)v.size() > cap) return fail(ctx, FE_ERR_CAPACITY, "output buffer too small");
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(fe_point_t));
  return FE_OK;
}

int fe_filter_cloud(fe_ctx_t* ctx, const fe_point_t* in, int64_t n, fe_point_t* out, int64_t cap, int64_t* n_out) {
  if (!ctx || n < 0 || (n > 0 && !in)) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (n_out) *n_out = 0;
  if (n == 0) return FE_OK;
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, true);
  if (st) return st;
  st = ensure_kc(ctx, s);
  if (st) return st;
  int nch;
  st = stage_upload_k1(ctx, s, in, n, F_CROP, &nch);
  if (st) return st;
  std::vector<int64_t> off;
  std::vector<fe_point_t> pts;
  st = gather_to_host(ctx, s, 1, true, s.d_crop, s.d_cropCnt, nullptr, off, pts);
  if (st) return st;
  return copy_points_out(ctx, pts, out, cap, n_out);
}

int fe_extract_clusters(fe_ctx_t* ctx, const fe_point_t* cloud, int64_t n, double tolerance,
                        int32_t min_size, int32_t max_size, int32_t* cluster_offsets,
                        int32_t cap_clusters, int32_t* indices, int64_t cap_indices, int32_t* n_clusters) {
  if (!ctx || n < 0 || (n > 0 && !cloud) || !n_clusters || !cluster_offsets) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  *n_clusters = 0;
  cluster_offsets[0] = 0;
  if (n == 0) return FE_OK;
  if (n > ECAP_G) return fail(ctx, FE_ERR_CAPACITY, "fe_extract_clusters: more than 65535 points");
  if (!(tolerance > 0.0)) return fail(ctx, FE_ERR_INVALID, "tolerance must be > 0");
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, true);
  if (st) return st;
  if (n > s.capPts) return fail(ctx, FE_ERR_CAPACITY, "fe_extract_clusters: more points than max_points_per_call");
  CK(cudaMemcpyAsync(s.d_pts, cloud, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, s.stream));
  const float tol_f = (float)tolerance;
  const float r2f = radius_sq_as_flann_sees_it((double)tol_f);
  int* d_off = (int*)s.d_keyA;
  int* d_idx = (int*)s.d_keyB;
  int* d_n = (int*)s.d_valA;
  if (n <= ECAP_L)
    k_extract_clusters_stage<false><<<1, NT2, kClusterSmemL, s.stream>>>(s.d_pts, (int)n, tol_f, r2f, min_size, max_size, d_off, (int)n,
                                                                         d_idx, d_n, nullptr);
  else
    k_extract_clusters_stage<true><<<1, NT2, cluster_smem_bytes_global(NT2), s.stream>>>(s.d_pts, (int)n, tol_f, r2f, min_size, max_size,
                                                                                          d_off, (int)n, d_idx, d_n, s.d_slabs);
  ctx->launches++;
  CK(cudaGetLastError());
  int nc = 0;
  CK(cudaMemcpyAsync(&nc, d_n, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  *n_clusters = nc;
  if (nc > cap_clusters) return fail(ctx, FE_ERR_CAPACITY, "cluster_offsets too small");
  std::vector<int> off(nc + 1);
  CK(cudaMemcpy(off.data(), d_off, (size_t)(nc + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  memcpy(cluster_offsets, off.data(), (size_t)(nc + 1) * sizeof(int));
  if (off[nc] > cap_indices) return fail(ctx, FE_ERR_CAPACITY, "indices too small");
  if (off[nc] > 0 && indices) CK(cudaMemcpy(indices, d_idx, (size_t)off[nc] * sizeof(int), cudaMemcpyDeviceToHost));
  return FE_OK;
}

static int stage_keypoints(fe_ctx* ctx, const fe_point_t* cloud, int64_t n, bool singleRing, bool merge,
                           fe_point_t* kp, int64_t capKp, int64_t* nKp, fe_point_t* kc, int64_t capKc, int64_t* nKc) {
  if (nKp) *nKp = 0;
  if (nKc) *nKc = 0;
  if (n == 0) return FE_OK;
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, true);
  if (st) return st;
  st = ensure_kc(ctx, s);
  if (st) return st;
  int nch;
  st = stage_upload_k1(ctx, s, cloud, n, singleRing ? 0 : F_RING, &nch);
  if (st) return st;
  CK(cudaMemsetAsync(s.d_ctr, 0, sizeof(DevCounters), s.stream));
  launch_clustering(ctx, s, 1, singleRing, true, merge);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (s.h_ctr->err) return fail(ctx, FE_ERR_CAPACITY, err_bits(s.h_ctr->err));
  std::vector<int64_t> off;
  std::vector<fe_point_t> pts;
  if (merge) {
    int base = 0, cnt = 0;
    CK(cudaMemcpy(&base, s.d_kpBase, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cnt, s.d_kpCnt, sizeof(int), cudaMemcpyDeviceToHost));
    pts.resize(cnt);
    if (cnt) CK(cudaMemcpy(pts.data(), s.d_kpPool + base, (size_t)cnt * sizeof(float4), cudaMemcpyDeviceToHost));
  } else {
    st = gather_to_host(ctx, s, 1, false, s.d_kfPool, s.d_kfCnt, s.d_kfBase, off, pts);
    if (st) return st;
  }
  st = copy_points_out(ctx, pts, kp, capKp, nKp);
  if (st) return st;
  st = gather_to_host(ctx, s, 1, false, s.d_kcPool, s.d_kcCnt, s.d_kcBase, off, pts);
  if (st) return st;
  return copy_points_out(ctx, pts, kc, capKc, nKc);
}

int fe_get_cylinder_segments(fe_ctx_t* ctx, const fe_point_t* ring_cloud, int64_t n, fe_point_t* centroids,
                             int64_t cap_centroids, int64_t* n_centroids, fe_point_t* cluster_cloud,
                             int64_t cap_cloud, int64_t* n_cloud) {
  if (!ctx || n < 0 || (n > 0 && !ring_cloud)) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return stage_keypoints(ctx, ring_cloud, n, true, false, centroids, cap_centroids, n_centroids, cluster_cloud, cap_cloud, n_cloud);
}

int fe_estimate_keypoints(fe_ctx_t* ctx, const fe_point_t* cloud, int64_t n, fe_point_t* keypoints,
                          int64_t cap_keypoints, int64_t* n_keypoints, fe_point_t* keypoint_cloud,
                          int64_t cap_cloud, int64_t* n_cloud) {
  if (!ctx || n < 0 || (n > 0 && !cloud)) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return stage_keypoints(ctx, cloud, n, false, true, keypoints, cap_keypoints, n_keypoints, keypoint_cloud, cap_cloud, n_cloud);
}

int fe_estimate_descriptors(fe_ctx_t* ctx, const fe_point_t* cloud_full, int64_t n, const fe_point_t* keypoints,
                            int64_t k, float* descriptors) {
  if (!ctx || n < 0 || k < 0 || (n > 0 && !cloud_full) || (k > 0 && (!keypoints || !descriptors))) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (k == 0) return FE_OK;  // src:331-332
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, true);
  if (st) return st;
  if (k > s.capKp) return fail(ctx, FE_ERR_CAPACITY, "more keypoints than max_keypoints_per_call");
  // the search surface only matters around the keypoints: grid over their bounding box
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < k; i++) {
    const float v[3] = {keypoints[i].x, keypoints[i].y, keypoints[i].z};
    if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) continue;
    for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], v[d]); hi[d] = std::max(hi[d], v[d]); }
  }
  if (!(lo[0] <= hi[0])) { for (int d = 0; d < 3; d++) { lo[d] = 0.f; hi[d] = 0.f; } }
  DevParams saved = ctx->dp;
  surface_grid(ctx->dp, ctx->params.descriptor_radius, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
  DevParams P = ctx->dp;
  auto restore = [&](int code) { ctx->dp = saved; return code; };
  int nch = 0;
  st = stage_upload_k1(ctx, s, cloud_full, n, F_SURF, &nch);
  if (st) return restore(st);
  if (n == 0) {  // one empty scan
    int64_t offs[2] = {0, 0}; int64_t np0;
    st = stage_scans(ctx, s, offs, nullptr, 1, &np0, &nch);
    if (st) return restore(st);
  }
  st = ensure_rowstart(ctx, s, 1);
  if (st) return restore(st);
  cudaStream_t q = s.stream;
  std::vector<int> kpOff = {0, (int)k};
  std::vector<int> kpScan((size_t)k, 0);
  if (cudaMemsetAsync(s.d_ctr, 0, sizeof(DevCounters), q) != cudaSuccess ||
      cudaMemcpyAsync(s.d_kpOff, kpOff.data(), 2 * sizeof(int), cudaMemcpyHostToDevice, q) != cudaSuccess ||
      cudaMemcpyAsync(s.d_kpScan, kpScan.data(), (size_t)k * sizeof(int), cudaMemcpyHostToDevice, q) != cudaSuccess ||
      cudaMemcpyAsync(s.d_kpOut, keypoints, (size_t)k * sizeof(float4), cudaMemcpyHostToDevice, q) != cudaSuccess)
    return restore(fail(ctx, FE_ERR_CUDA, "staging of the descriptor inputs failed"));
  launch_surface_grid(ctx, s, 1, P, s.stream);
  launch_desc_mark(ctx, s, 1, P, ctx->numSms * 4);
  launch_density(ctx, s, 1, P, n);
  launch_desc_hist(ctx, s, 1, P, ctx->numSms * 4, false);
  ctx->dp = saved;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, q));
  CK(cudaMemcpyAsync(descriptors, s.d_desc, (size_t)k * FE_DESC_LEN * sizeof(float), cudaMemcpyDeviceToHost, q));
  CK(cudaStreamSynchronize(q));
  if (s.h_ctr->err) return fail(ctx, FE_ERR_CAPACITY, err_bits(s.h_ctr->err));
  return FE_OK;
}

int fe_pack_point_descriptors(const fe_point_t* keypoints, const float* descriptors, int64_t k, float* records) {
  if (k < 0 || (k > 0 && (!keypoints || !descriptors || !records))) return FE_ERR_INVALID;
  for (int64_t i = 0; i < k; i++) {
    float* r = records + i * FE_RECORD_FLOATS;
    memset(r, 0, FE_RECORD_FLOATS * sizeof(float));
    // PCL_ADD_POINT4D: x,y,z + one pad float (not a registered field: stays value-initialised 0)
    r[0] = keypoints[i].x; r[1] = keypoints[i].y; r[2] = keypoints[i].z;
    r[4] = keypoints[i].intensity;
    memcpy(r + 5, descriptors + i * FE_DESC_LEN, FE_DESC_LEN * sizeof(float));
    // rf[9] at r + 5 + 1980 stays zero (3dsc.hpp zeroes it); 2 floats of tail padding (EIGEN_ALIGN16)
  }
  return FE_OK;
}

// ---- scan-parallel sharding over several GPUs in one process -----------------------------------------
struct fe_multi {
  std::vector<fe_ctx_t*> ctx;
  std::vector<int64_t> kpOffsets;
  bool cloudOutputs = false;
  std::vector<int64_t> cloudOff, kcOff;
  std::vector<fe_point_t> cloudPts, kcPts;
  fe_point_t* kp = nullptr;   // gathered results: plain malloc'd buffers grown on demand (no zero fill)
  float* desc = nullptr;
  int64_t capKp = 0, capDesc = 0;
  std::string err;
};

int fe_multi_create(const int32_t* devices, int32_t n_devices, const fe_params_t* params,
                    const fe_limits_t* limits, fe_multi_t** out) {
  if (!out || !devices || n_devices < 1 || !params) return FE_ERR_INVALID;
  *out = nullptr;
  fe_multi* m = new fe_multi();
  for (int g = 0; g < n_devices; g++) {
    fe_ctx_t* c = nullptr;
    const int st = fe_create(devices[g], params, limits, &c);
    if (st != FE_OK) { fe_multi_destroy(m); return st; }
    m->ctx.push_back(c);
  }
  *out = m;
  return FE_OK;
}

int fe_multi_enable_record_output(fe_multi_t* m, int32_t enable) {
  if (!m) return FE_ERR_INVALID;
  for (fe_ctx_t* c : m->ctx) {
    const int st = fe_enable_record_output(c, enable);
    if (st) return st;
  }
  return FE_OK;
}

int fe_multi_enable_cloud_outputs(fe_multi_t* m, int32_t enable) {
  if (!m) return FE_ERR_INVALID;
  for (fe_ctx_t* c : m->ctx) {
    const int st = fe_enable_cloud_outputs(c, enable);
    if (st) return st;
  }
  m->cloudOutputs = enable != 0;
  return FE_OK;
}

int fe_multi_get_cloud_outputs(fe_multi_t* m, const int64_t** cloud_offsets, const fe_point_t** cloud,
                               const int64_t** kpcloud_offsets, const fe_point_t** keypoint_cloud) {
  if (!m) return FE_ERR_INVALID;
  if (m->cloudOff.empty()) { m->err = "no cloud outputs recorded (fe_multi_enable_cloud_outputs before fe_multi_process_batch)"; return FE_ERR_INVALID; }
  if (cloud_offsets) *cloud_offsets = m->cloudOff.data();
  if (cloud) *cloud = m->cloudPts.data();
  if (kpcloud_offsets) *kpcloud_offsets = m->kcOff.data();
  if (keypoint_cloud) *keypoint_cloud = m->kcPts.data();
  return FE_OK;
}

void fe_multi_destroy(fe_multi_t* m) {
  if (!m) return;
  for (fe_ctx_t* c : m->ctx) fe_destroy(c);
  free(m->kp);
  free(m->desc);
  delete m;
}

const char* fe_multi_last_error(const fe_multi_t* m) { return m ? m->err.c_str() : "null context"; }

int fe_multi_process_batch(fe_multi_t* m, const fe_point_t* points, const int64_t* scan_offsets,
                           const double* roll_pitch, int32_t n_scans, fe_batch_result_t* out) {
  if (!m || !out || n_scans < 0 || (n_scans > 0 && (!scan_offsets || !roll_pitch))) return FE_ERR_INVALID;
  const int G = (int)m->ctx.size();
  m->err.clear();
  std::vector<fe_batch_result_t> res(G);
  std::vector<int> status(G, FE_OK);
  std::vector<int> lo(G + 1);
  for (int g = 0; g <= G; g++) lo[g] = (int)(((int64_t)n_scans * g) / G);
  {  // one host thread per GPU: its shard through fe_process_batch (absolute offsets into `points`)
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
      th.emplace_back([&, g]() {
        status[g] = fe_process_batch(m->ctx[g], points, scan_offsets + lo[g], roll_pitch ? roll_pitch + 2 * (int64_t)lo[g] : nullptr,
                                     lo[g + 1] - lo[g], &res[g]);
      });
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < G; g++)
    if (status[g] != FE_OK) {
      m->err = std::string("device shard ") + std::to_string(g) + ": " + fe_last_error(m->ctx[g]);
      return status[g];
    }
  // host-side gather in scan order
  std::vector<int64_t> kbase(G + 1, 0);
  for (int g = 0; g < G; g++) kbase[g + 1] = kbase[g] + res[g].n_keypoints;
  const int64_t K = kbase[G];
  const bool desc = m->ctx[0]->params.estimate_descriptors != 0;
  const size_t dl = m->ctx[0]->recordOutput ? FE_RECORD_FLOATS : FE_DESC_LEN;
  m->kpOffsets.assign((size_t)n_scans + 1, 0);
  if (K > m->capKp) {
    free(m->kp);
    m->capKp = K + K / 2 + 1024;
    m->kp = (fe_point_t*)malloc((size_t)m->capKp * sizeof(fe_point_t));
    if (!m->kp) { m->capKp = 0; m->err = "out of host memory"; return FE_ERR_CAPACITY; }
  }
  if (desc && K > m->capDesc) {
    free(m->desc);
    m->capDesc = K + K / 2 + 1024;
    m->desc = (float*)malloc((size_t)m->capDesc * FE_RECORD_FLOATS * sizeof(float));
    if (!m->desc) { m->capDesc = 0; m->err = "out of host memory"; return FE_ERR_CAPACITY; }
  }
  {
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
      th.emplace_back([&, g]() {
        const int ns = lo[g + 1] - lo[g];
        for (int i = 0; i <= ns; i++) m->kpOffsets[lo[g] + i] = kbase[g] + res[g].keypoint_offsets[i];
        if (res[g].n_keypoints > 0) {
          memcpy(m->kp + kbase[g], res[g].keypoints, (size_t)res[g].n_keypoints * sizeof(fe_point_t));
          if (desc) memcpy(m->desc + kbase[g] * dl, res[g].descriptors, (size_t)res[g].n_keypoints * dl * sizeof(float));
        }
      });
    for (auto& t : th) t.join();
  }
  m->cloudOff.clear(); m->kcOff.clear();
  if (m->cloudOutputs) {  // the shards' ~cloud / ~keypoint_cloud concatenated in scan order
    m->cloudOff.assign((size_t)n_scans + 1, 0); m->kcOff.assign((size_t)n_scans + 1, 0);
    m->cloudPts.clear(); m->kcPts.clear();
    for (int g = 0; g < G; g++) {
      const int ns = lo[g + 1] - lo[g];
      if (ns == 0) continue;
      const int64_t *co = nullptr, *ko = nullptr;
      const fe_point_t *cp = nullptr, *kp2 = nullptr;
      const int st = fe_get_cloud_outputs(m->ctx[g], &co, &cp, &ko, &kp2);
      if (st != FE_OK) { m->err = std::string("device shard ") + std::to_string(g) + ": " + fe_last_error(m->ctx[g]); return st; }
      const int64_t cb = (int64_t)m->cloudPts.size(), kb = (int64_t)m->kcPts.size();
      for (int i = 0; i <= ns; i++) { m->cloudOff[lo[g] + i] = cb + co[i]; m->kcOff[lo[g] + i] = kb + ko[i]; }
      m->cloudPts.insert(m->cloudPts.end(), cp, cp + co[ns]);
      m->kcPts.insert(m->kcPts.end(), kp2, kp2 + ko[ns]);
    }
  }
  out->n_scans = n_scans;
  out->n_keypoints = K;
  out->keypoint_offsets = m->kpOffsets.data();
  out->keypoints = m->kp;
  out->descriptors = desc ? m->desc : nullptr;
  out->on_device = 0;
  out->gpu_launches = 0;
  for (int g = 0; g < G; g++) out->gpu_launches += res[g].gpu_launches;
  return FE_OK;
}

// bench hook: work of the 3DSC density stage of the last device call: out = {distance tests, marked surface points
// (distinct points inside some keypoint's support sphere), halo points kept by the cell grid}
int fe_debug_density_work(fe_ctx_t* ctx, int64_t out[3]) {
  if (!ctx || !out || !ctx->slot[0].stream) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  CK(cudaStreamSynchronize(s.stream));
  out[0] = (int64_t)s.h_ctr->dens_work[0];
  out[1] = (int64_t)s.h_ctr->dens_work[1];
  out[2] = 0;
  if (s.nscans > 0 && ctx->params.estimate_descriptors) {
    std::vector<int> sn((size_t)s.nscans);
    CK(cudaMemcpy(sn.data(), s.d_surfN, sn.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int v : sn) out[2] += v;
  }
  return FE_OK;
}

// bench hook: bare pinned-host -> device copies of `bytes` from `host` into the context's staging buffer, chunk after
// chunk on one stream, timed with CUDA events — the ceiling the host entry points can reach on this box
int fe_debug_h2d_probe(fe_ctx_t* ctx, const void* host, int64_t bytes, float* ms) {
  if (!ctx || !host || bytes < 0 || !ms) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  int st = ensure_slot(ctx, s, true);
  if (st) return st;
  const int64_t chunk = s.capPts * (int64_t)sizeof(float4);
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaEventRecord(s.evT0, s.stream));
  for (int64_t off = 0; off < bytes; off += chunk)
    CK(cudaMemcpyAsync(s.d_pts, (const char*)host + off, (size_t)std::min(chunk, bytes - off), cudaMemcpyHostToDevice, s.stream));
  CK(cudaEventRecord(s.evT1, s.stream));
  CK(cudaEventSynchronize(s.evT1));
  CK(cudaEventElapsedTime(ms, s.evT0, s.evT1));
  return FE_OK;
}

// debug / test hook: run K2 through the grid-based kernels only (the general fallback of the run-based kernel), so
// that tests can cross-check the two algorithms against each other and against the oracle
int fe_debug_force_grid_clustering(fe_ctx_t* ctx, int32_t enable) {
  if (!ctx) return FE_ERR_INVALID;
  for (int k = 0; k < 2; k++) if (ctx->slot[k].stream) CK(cudaStreamSynchronize(ctx->slot[k].stream));
  ctx->gridClustering = enable != 0;
  ctx->epoch++;
  return FE_OK;
}

// debug / test hook: switch the CUDA-graph replay of small sub-batches off (eager launches) or on; out[0] (nullable)
// receives the number of graph replays so far
int fe_debug_enable_graphs(fe_ctx_t* ctx, int32_t enable, int64_t* replays) {
  if (!ctx) return FE_ERR_INVALID;
  ctx->useGraphs = enable != 0;
  if (replays) *replays = ctx->graphReplays;
  return FE_OK;
}

// debug / test hook: how many small sub-batches had to be run again with the whole chain of fallback kernels
int fe_debug_lean_reruns(fe_ctx_t* ctx, int64_t* reruns) {
  if (!ctx || !reruns) return FE_ERR_INVALID;
  *reruns = ctx->leanReruns;
  return FE_OK;
}

// debug / test hook: keypoints_full (src:205), i.e. estimateKeypoints before the cross-ring merge
int fe_debug_keypoints_full(fe_ctx_t* ctx, const fe_point_t* cloud, int64_t n, fe_point_t* kf, int64_t cap, int64_t* n_kf) {
  if (!ctx || n < 0 || (n > 0 && !cloud)) return FE_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return stage_keypoints(ctx, cloud, n, false, false, kf, cap, n_kf, nullptr, 0, nullptr);
}

// debug / test hook: the cluster order PCL's final std::sort leaves, from the replay the kernels use
void fe_debug_sort_replay(const int32_t* sizes, int32_t n, int32_t* order_out) {
  std::vector<int> ids(n);
  for (int i = 0; i < n; i++) ids[i] = i;
  fe::pcl_cluster_order(ids.data(), n, [=](int id) { return sizes[id]; });
  for (int i = 0; i < n; i++) order_out[i] = ids[i];
}

}  // extern "C"

// fe_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: nothing in the product path may link,
// import or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / the CPU arm.
//
// PARITY UNPINNED.  The reference (GAVLab/feature_extraction) ships no tests, fixtures or golden
// vectors (CMakeLists.txt:44-51 is commented-out boilerplate) and its arithmetic lives in
// un-vendored third parties that are absent from /root/reference and from this image:
//   PCL >= 1.8.0 (package.xml:52,62), Eigen 3.2.x, FLANN 1.8.x, Boost (mt19937/uniform_01).
// This file restates (1) src/feature_extraction_node.cpp:147-355 literally and (2) the published
// algorithms of the PCL 1.8.0 / Eigen 3.2 / FLANN 1.8 functions it calls, each as a named,
// isolated function so a disagreement with a real PCL build can be fixed in one place.  The only
// hard pins are the known-answer constants of SURVEY.md §8(c) (tests/test_oracle_pins.py).
//
// Build: g++ -O2 -ffp-contract=off (no -march=native, no -ffast-math): unfused float arithmetic
// is what a stock x86-64 PCL build executes.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

#include "../include/fe_b200.h"  // types only (fe_point_t, fe_params_t)

namespace {

typedef fe_point_t P4;

// ------------------------------------------------------------------------------------------
// Row A — getElevationAngles, src:147-156.  double throughout, stored as float.
// ------------------------------------------------------------------------------------------
void get_elevation_angles(P4* pts, int64_t n) {
  double x, xp, y, z, az, el_deg;
  for (int64_t i = 0; i < n; i++) {
    x = pts[i].x; y = pts[i].y; z = pts[i].z;          // src:150
    az = atan2(y, x);                                  // src:151
    xp = cos(az) * x + sin(az) * y;                    // src:152
    el_deg = atan2(z, xp) * 180 / M_PI;                // src:153
    pts[i].intensity = (float)el_deg;                  // src:154
  }
}

// ------------------------------------------------------------------------------------------
// Row B — rotateCloud, src:159-167.
// Eigen 3.2: AngleAxisf(a, axis) -> Quaternionf {w = cos(a/2), vec = sin(a/2)*axis};
// AngleAxis*AngleAxis = Quaternion product; Transform::rotate(q) = linear * q.toRotationMatrix().
// ------------------------------------------------------------------------------------------
struct Quatf { float w, x, y, z; };

Quatf eigen_quat_from_angle_axis(float angle, float ax, float ay, float az) {
  float ha = 0.5f * angle;
  float s = sinf(ha);
  Quatf q;
  q.w = cosf(ha);
  q.x = s * ax; q.y = s * ay; q.z = s * az;
  return q;
}

Quatf eigen_quat_mul(const Quatf& a, const Quatf& b) {
  Quatf r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

void eigen_quat_to_matrix(const Quatf& q, float R[3][3]) {
  const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
  const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1.0f - (tyy + tzz); R[0][1] = txy - twz;          R[0][2] = txz + twy;
  R[1][0] = txy + twz;          R[1][1] = 1.0f - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;          R[2][1] = tyz + twx;          R[2][2] = 1.0f - (txx + tyy);
}

// Eigen's unrolled, non-vectorised reduction of three terms is a0 + (a1 + a2)
// (redux_novec_unroller splits Length 3 as 1 | 2).
inline float eigen_sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

void rotation_matrix(double roll, double pitch, float m[9]) {
  // src:163-164: AngleAxisf(pitch, UnitY) * AngleAxisf(roll, UnitX); doubles narrow to float.
  Quatf qy = eigen_quat_from_angle_axis((float)pitch, 0.0f, 1.0f, 0.0f);
  Quatf qx = eigen_quat_from_angle_axis((float)roll, 1.0f, 0.0f, 0.0f);
  Quatf q = eigen_quat_mul(qy, qx);
  float R[3][3];
  eigen_quat_to_matrix(q, R);
  // Affine3f::Identity().rotate(q): linear = Identity * R (coefficient-wise 3x3 product)
  const float I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      m[i * 3 + j] = eigen_sum3(I[i][0] * R[0][j], I[i][1] * R[1][j], I[i][2] * R[2][j]);
}

// pcl::transformPointCloud, PCL 1.8.0 scalar form (common/impl/transforms.hpp):
//   x' = (float)(m00*x + m01*y + m02*z + m03), left to right in float; translation is 0 (src:162)
void rotate_cloud(P4* pts, int64_t n, double roll, double pitch) {
  float m[9];
  rotation_matrix(roll, pitch, m);
  const float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
  for (int64_t i = 0; i < n; i++) {
    const float x = pts[i].x, y = pts[i].y, z = pts[i].z;
    pts[i].x = m[0] * x + m[1] * y + m[2] * z + t0;
    pts[i].y = m[3] * x + m[4] * y + m[5] * z + t1;
    pts[i].z = m[6] * x + m[7] * y + m[8] * z + t2;
  }
}

// ------------------------------------------------------------------------------------------
// Row D / E1 — pcl::PassThrough (filters/impl/passthrough.hpp): limits stored as float,
// inclusive both ends, non-finite xyz or field value dropped, stable.
// ------------------------------------------------------------------------------------------
enum Field { FX = 0, FY = 1, FZ = 2, FI = 3 };

inline float field_of(const P4& p, Field f) {
  return f == FX ? p.x : f == FY ? p.y : f == FZ ? p.z : p.intensity;
}

void pass_through(const std::vector<P4>& in, Field f, double lo_d, double hi_d,
                  std::vector<P4>& out, std::vector<int>* kept_idx = nullptr) {
  const float lo = (float)lo_d, hi = (float)hi_d;  // setFilterLimits(const float&, const float&)
  std::vector<P4> tmp;
  tmp.reserve(in.size());
  if (kept_idx) kept_idx->clear();
  for (size_t i = 0; i < in.size(); i++) {
    const P4& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const float v = field_of(p, f);
    if (!std::isfinite(v)) continue;
    if (v < lo || v > hi) continue;
    tmp.push_back(p);
    if (kept_idx) kept_idx->push_back((int)i);
  }
  out.swap(tmp);
}

void filter_cloud(const fe_params_t& P, std::vector<P4>& cloud) {  // src:169-183
  pass_through(cloud, FZ, P.z_min, P.z_max, cloud);
  pass_through(cloud, FY, P.y_min, P.y_max, cloud);
  pass_through(cloud, FX, P.x_min, P.x_max, cloud);
}

// ------------------------------------------------------------------------------------------
// Neighbour search — FLANN L2_Simple<float> distance, strict d2 < r2 (RadiusResultSet).
// Two back ends with the same predicate: brute force (ground truth) and a leaf-15 KD-tree
// (PCL's algorithmic shape, used when the oracle is timed as the CPU baseline).
// ------------------------------------------------------------------------------------------
inline float l2_simple(const P4& a, const P4& b) {
  float result = 0.0f;
  float diff;
  diff = a.x - b.x; result += diff * diff;
  diff = a.y - b.y; result += diff * diff;
  diff = a.z - b.z; result += diff * diff;
  return result;
}

// KdTreeFLANN::radiusSearch: static_cast<float>(radius * radius), radius a double.
inline float radius_sq_float(double radius) { return (float)(radius * radius); }

// Tolerance-boundary report (BASELINE.json north_star: "points lying within 1e-6 m of a tolerance
// boundary reported separately").  Every radius predicate of the path is d2 < r2f in float; a pair is
// "on the boundary" when |sqrt((double)d2) - sqrt((double)r2f)| < eps.  Slots: 0 ring clustering
// (src:269-276), 1 cross-ring merge (src:222-229), 2 3DSC support radius (src:350), 3 3DSC point
// density radius (src:352).  The product counts the same pairs with its own kernels
// (fe_enable_boundary_report); the two must agree.
struct BoundaryCounter {
  double eps = 0.0;
  int64_t n[4] = {0, 0, 0, 0};
};

struct Hit { float d2; int idx; };
inline bool hit_less(const Hit& a, const Hit& b) {  // flann::DistanceIndex::operator<
  return (a.d2 < b.d2) || ((a.d2 == b.d2) && a.idx < b.idx);
}

struct Searcher {
  const P4* pts = nullptr;
  int n = 0;
  bool use_tree = false;
  bool sorted = false;
  // kd-tree
  struct Node { float lo[3], hi[3]; int left, right, begin, end; };
  std::vector<Node> nodes;
  std::vector<int> order;
  // boundary report: pairs (query, point) evaluated by radius() whose distance is within eps of the radius
  double bnd_eps = 0.0;
  mutable int64_t bnd_hits = 0;

  void set_input(const P4* p, int n_, bool tree, bool sorted_) {
    pts = p; n = n_; use_tree = tree; sorted = sorted_;
    nodes.clear(); order.clear();
    if (use_tree && n > 0) {
      order.resize(n);
      int m = 0;
      for (int i = 0; i < n; i++)  // KdTreeFLANN::convertCloudToArray skips non-finite points
        if (std::isfinite(p[i].x) && std::isfinite(p[i].y) && std::isfinite(p[i].z)) order[m++] = i;
      order.resize(m);
      if (m > 0) { nodes.reserve(2 * (m / 8) + 4); build(0, m); }
    }
  }
  static inline float coord(const P4& p, int d) { return d == 0 ? p.x : d == 1 ? p.y : p.z; }
  int build(int b, int e) {
    Node nd;
    for (int d = 0; d < 3; d++) { nd.lo[d] = std::numeric_limits<float>::infinity(); nd.hi[d] = -nd.lo[d]; }
    for (int i = b; i < e; i++)
      for (int d = 0; d < 3; d++) {
        float v = coord(pts[order[i]], d);
        nd.lo[d] = std::min(nd.lo[d], v); nd.hi[d] = std::max(nd.hi[d], v);
      }
    nd.left = nd.right = -1; nd.begin = b; nd.end = e;
    int id = (int)nodes.size();
    nodes.push_back(nd);
    if (e - b > 15) {  // FLANN KDTreeSingleIndex leaf_max_size = 15 (PCL default)
      int dim = 0;
      float best = nd.hi[0] - nd.lo[0];
      for (int d = 1; d < 3; d++) if (nd.hi[d] - nd.lo[d] > best) { best = nd.hi[d] - nd.lo[d]; dim = d; }
      float split = 0.5f * (nd.lo[dim] + nd.hi[dim]);
      int* first = order.data() + b;
      int* last = order.data() + e;
      int* mid = std::partition(first, last, [&](int i) { return coord(pts[i], dim) < split; });
      if (mid == first || mid == last) {
        mid = first + (e - b) / 2;
        std::nth_element(first, mid, last, [&](int a, int c) { return coord(pts[a], dim) < coord(pts[c], dim); });
      }
      int m = (int)(mid - order.data());
      int l = build(b, m);
      int r = build(m, e);
      nodes[id].left = l; nodes[id].right = r;
    }
    return id;
  }
  // all j with l2_simple(q, pts[j]) < r2f; returns count
  int radius(const P4& q, float r2f, std::vector<Hit>& out) const {
    out.clear();
    const bool bnd = bnd_eps > 0.0;
    const double bdist = bnd ? sqrt((double)r2f) : 0.0;
    const double r2prune = bnd ? (bdist + 2.0 * bnd_eps) * (bdist + 2.0 * bnd_eps) : (double)r2f;
    if (!use_tree) {
      for (int j = 0; j < n; j++) {
        float d2 = l2_simple(q, pts[j]);
        if (d2 < r2f) out.push_back({d2, j});
        if (bnd && fabs(sqrt((double)d2) - bdist) < bnd_eps) bnd_hits++;
      }
    } else if (!nodes.empty()) {
      int stack[128]; int sp = 0; stack[sp++] = 0;
      while (sp) {
        const Node& nd = nodes[stack[--sp]];
        double lb = 0.0;  // exact (double) lower bound of the box distance, pruned conservatively
        for (int d = 0; d < 3; d++) {
          double v = coord(q, d), t = 0.0;
          if (v < nd.lo[d]) t = (double)nd.lo[d] - v; else if (v > nd.hi[d]) t = v - (double)nd.hi[d];
          lb += t * t;
        }
        if (lb * (1.0 - 1e-6) >= r2prune) continue;
        if (nd.left < 0) {
          for (int i = nd.begin; i < nd.end; i++) {
            int j = order[i];
            float d2 = l2_simple(q, pts[j]);
            if (d2 < r2f) out.push_back({d2, j});
            if (bnd && fabs(sqrt((double)d2) - bdist) < bnd_eps) bnd_hits++;
          }
        } else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
      }
    }
    if (sorted) std::sort(out.begin(), out.end(), hit_less);
    else if (use_tree) std::sort(out.begin(), out.end(), [](const Hit& a, const Hit& b) { return a.idx < b.idx; });
    return (int)out.size();
  }
};

// ------------------------------------------------------------------------------------------
// Row E2 — pcl::EuclideanClusterExtraction::extract (segmentation/impl/extract_clusters.hpp):
// BFS per unprocessed point in index order, size gate after the full BFS, indices sorted
// ascending, clusters finally std::sort(rbegin, rend, size <).
// ------------------------------------------------------------------------------------------
typedef std::vector<std::vector<int> > Clusters;

bool compare_point_clusters(const std::vector<int>& a, const std::vector<int>& b) {
  return a.size() < b.size();
}

void euclidean_cluster_extract(const std::vector<P4>& cloud, double tolerance_d, int min_size,
                               int max_size, bool use_tree, Clusters& clusters,
                               BoundaryCounter* bc = nullptr, int bslot = 0) {
  clusters.clear();
  const int n = (int)cloud.size();
  if (n == 0) return;
  const float tolerance = (float)tolerance_d;           // extractEuclideanClusters(.., float tolerance, ..)
  const float r2f = radius_sq_float((double)tolerance);  // radiusSearch(.., double radius, ..)
  Searcher tree;
  tree.set_input(cloud.data(), n, use_tree, /*sorted=*/false);  // search::KdTree<PointT>(false)
  if (bc) tree.bnd_eps = bc->eps;
  std::vector<char> processed(n, 0);
  std::vector<Hit> nn;
  std::vector<int> seed_queue;
  for (int i = 0; i < n; i++) {
    if (processed[i]) continue;
    seed_queue.clear();
    size_t sq_idx = 0;
    seed_queue.push_back(i);
    processed[i] = 1;
    while (sq_idx < seed_queue.size()) {
      if (!tree.radius(cloud[seed_queue[sq_idx]], r2f, nn)) { sq_idx++; continue; }
      for (size_t j = 0; j < nn.size(); j++) {
        if (processed[nn[j].idx]) continue;
        seed_queue.push_back(nn[j].idx);
        processed[nn[j].idx] = 1;
      }
      sq_idx++;
    }
    if ((int)seed_queue.size() >= min_size && (int)seed_queue.size() <= max_size) {
      std::vector<int> r(seed_queue);
      std::sort(r.begin(), r.end());
      r.erase(std::unique(r.begin(), r.end()), r.end());
      clusters.push_back(r);
    }
  }
  std::sort(clusters.rbegin(), clusters.rend(), compare_point_clusters);
  // every point is the query of exactly one search, so each unordered pair was seen twice
  if (bc) bc->n[bslot] += tree.bnd_hits / 2;
}

// ------------------------------------------------------------------------------------------
// Row F — getCylinderSegments, src:261-327.
// ------------------------------------------------------------------------------------------
void get_cylinder_segments(const fe_params_t& P, const std::vector<P4>& cloud, bool use_tree,
                           std::vector<P4>& keypoints, std::vector<P4>& keypoint_cloud,
                           Clusters* clusters_out = nullptr, BoundaryCounter* bc = nullptr) {
  if (cloud.size() <= 0) return;  // src:263-264
  Clusters clusterIndices;
  euclidean_cluster_extract(cloud, P.cluster_tolerance, P.cluster_min_count, P.cluster_max_count,
                            use_tree, clusterIndices, bc, 0);  // src:269-276
  if (clusters_out) *clusters_out = clusterIndices;
  if (clusterIndices.size() <= 0) return;  // src:278-279
  for (size_t i = 0; i < clusterIndices.size(); ++i) {
    P4 pt_centroid;
    std::vector<P4> cluster;
    double x, y, z;
    double sumx = 0.0, sumy = 0.0, sumz = 0.0;
    double minx = 1000.0, maxx = -1000.0;  // src:289
    double miny = 1000.0, maxy = -1000.0;  // src:290
    int clusterSize = (int)clusterIndices[i].size();
    for (int j = 0; j < clusterSize; ++j) {
      const P4& s = cloud[clusterIndices[i][j]];
      x = s.x; y = s.y; z = s.z;
      sumx += x; sumy += y; sumz += z;
      if (x < minx) minx = x;
      if (y < miny) miny = y;
      if (x > maxx) maxx = x;
      if (y > maxy) maxy = y;
      P4 pt;
      pt.x = (float)x; pt.y = (float)y; pt.z = (float)z; pt.intensity = s.intensity;
      cluster.push_back(pt);
    }
    double diameter = pow(pow(maxx - minx, 2) + pow(maxy - miny, 2), 0.5);  // src:314
    if (diameter < (2 * P.cluster_radius_threshold)) {                      // src:316
      pt_centroid.x = (float)(sumx / ((double)clusterSize));
      pt_centroid.y = (float)(sumy / ((double)clusterSize));
      pt_centroid.z = (float)(sumz / ((double)clusterSize));
      pt_centroid.intensity = cloud[clusterIndices[i][0]].intensity;  // src:320
      keypoints.push_back(pt_centroid);
      keypoint_cloud.insert(keypoint_cloud.end(), cluster.begin(), cluster.end());  // src:323
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row E1 + G — estimateKeypoints, src:185-259.
// ------------------------------------------------------------------------------------------
void estimate_keypoints(const fe_params_t& P, const std::vector<P4>& cloud, bool use_tree,
                        std::vector<P4>& keypoints, std::vector<P4>& keypoint_cloud,
                        std::vector<P4>* keypoints_full_out = nullptr, BoundaryCounter* bc = nullptr) {
  std::vector<P4> keypoints_full;
  double channelElevationDegrees;
  for (int i = 0; i < 16; ++i) {  // src:195
    std::vector<P4> cylinderCentroids, cylinderCloud, channel;
    channelElevationDegrees = (i - 7) * 2 - 1;  // src:200
    pass_through(cloud, FI, channelElevationDegrees - 1.0, channelElevationDegrees + 1.0, channel);
    get_cylinder_segments(P, channel, use_tree, cylinderCentroids, cylinderCloud, nullptr, bc);
    keypoints_full.insert(keypoints_full.end(), cylinderCentroids.begin(), cylinderCentroids.end());
    keypoint_cloud.insert(keypoint_cloud.end(), cylinderCloud.begin(), cylinderCloud.end());
  }
  if (keypoints_full_out) *keypoints_full_out = keypoints_full;
  if (keypoints_full.size() <= 0) return;  // src:209-210

  std::vector<double> zhold(keypoints_full.size());  // src:213 (a stack VLA in the reference)
  for (size_t i = 0; i < keypoints_full.size(); ++i) {
    zhold[i] = keypoints_full[i].z;
    keypoints_full[i].z = (float)(keypoints_full[i].intensity * 0.75 * P.cluster_radius_threshold / 2);  // src:217
  }
  Clusters clusterIndices;
  euclidean_cluster_extract(keypoints_full, P.cluster_radius_threshold, P.number_detection_channels,
                            16, use_tree, clusterIndices, bc, 1);  // src:222-229
  for (size_t i = 0; i < keypoints_full.size(); ++i) keypoints_full[i].z = (float)zhold[i];  // src:231-232
  if (clusterIndices.size() <= 0) return;  // src:234-235

  for (size_t i = 0; i < clusterIndices.size(); ++i) {
    P4 pt_centroid;
    double sumx = 0.0, sumy = 0.0, sumz = 0.0;
    int clusterSize = (int)clusterIndices[i].size();
    for (int j = 0; j < clusterSize; ++j) {
      sumx += keypoints_full[clusterIndices[i][j]].x;
      sumy += keypoints_full[clusterIndices[i][j]].y;
      sumz += keypoints_full[clusterIndices[i][j]].z;
    }
    pt_centroid.x = (float)(sumx / ((double)clusterSize));
    pt_centroid.y = (float)(sumy / ((double)clusterSize));
    pt_centroid.z = (float)(sumz / ((double)clusterSize));
    pt_centroid.intensity = keypoints_full[clusterIndices[i][0]].intensity;  // src:254
    keypoints.push_back(pt_centroid);
  }
}

// ------------------------------------------------------------------------------------------
// Rows H-N — pcl::ShapeContext3DEstimation (features/impl/3dsc.hpp, PCL 1.8.0).
// ------------------------------------------------------------------------------------------
struct ShapeContextTables {
  float radii[16];     // radii_interval_
  float theta[12];     // theta_divisions_
  float phi[13];       // phi_divisions_
  float lut[FE_DESC_LEN];  // volume_lut_
};

inline float pcl_deg2rad(float a) { return a * 0.017453293f; }
inline float pcl_rad2deg(float a) { return a * 57.29578f; }

void sc3d_init_compute(double search_radius, double min_radius, ShapeContextTables& T) {
  const size_t azimuth_bins = 12, elevation_bins = 11, radius_bins = 15;
  float azimuth_interval = 360.0f / static_cast<float>(azimuth_bins);
  float elevation_interval = 180.0f / static_cast<float>(elevation_bins);
  for (size_t j = 0; j < radius_bins + 1; j++)
    T.radii[j] = static_cast<float>(exp(log(min_radius) + ((static_cast<float>(j) / static_cast<float>(radius_bins)) * log(search_radius / min_radius))));
  for (size_t k = 0; k < elevation_bins + 1; k++) T.theta[k] = static_cast<float>(k) * elevation_interval;
  for (size_t l = 0; l < azimuth_bins + 1; l++) T.phi[l] = static_cast<float>(l) * azimuth_interval;
  float integr_phi = pcl_deg2rad(T.phi[1]) - pcl_deg2rad(T.phi[0]);
  float e = 1.0f / 3.0f;
  for (size_t j = 0; j < radius_bins; j++) {
    float integr_r = (T.radii[j + 1] * T.radii[j + 1] * T.radii[j + 1] / 3.0f) - (T.radii[j] * T.radii[j] * T.radii[j] / 3.0f);
    for (size_t k = 0; k < elevation_bins; k++) {
      float integr_theta = cosf(pcl_deg2rad(T.theta[k])) - cosf(pcl_deg2rad(T.theta[k + 1]));
      float V = integr_phi * integr_theta * integr_r;
      for (size_t l = 0; l < azimuth_bins; l++)
        T.lut[(l * elevation_bins * radius_bins) + k * radius_bins + j] = 1.0f / powf(V, e);
    }
  }
}

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 sub(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float eigen_dot(const V3& a, const V3& b) { return eigen_sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float eigen_norm(const V3& a) { return sqrtf(eigen_dot(a, a)); }
// Eigen 3.2 `v /= s` on a floating-point vector multiplies by Scalar(1)/s; normalize() is
// `*this /= norm()` with no zero check (a zero vector becomes NaN).
inline void eigen32_normalize(V3& a) {
  const float inv = 1.0f / eigen_norm(a);
  a.x *= inv; a.y *= inv; a.z *= inv;
}
inline V3 eigen_cross(const V3& a, const V3& b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline bool pcl_utils_equal(float a, float b) {  // pcl::utils::equal, eps = numeric_limits<float>::min()
  return fabsf(a - b) < std::numeric_limits<float>::min();
}

struct DescDebug {      // per keypoint, optional
  float edge_margin;    // min over neighbours of the relative distance of theta/phi to a bin edge
  int n_neighbors;
};

// estimateDescriptors, src:329-355.  `desc` is K x 1980.
void estimate_descriptors(const fe_params_t& P, const std::vector<P4>& cloud,
                          const std::vector<P4>& keypoints, bool use_tree, float* desc,
                          DescDebug* dbg = nullptr, BoundaryCounter* bc = nullptr) {
  if (keypoints.size() <= 0) return;  // src:331-332
  const V3 normal_const = v3(0.0f, 0.0f, 1.0f);  // src:337-340, every surface normal
  const double search_radius = P.descriptor_radius;          // src:350
  const double min_radius = P.descriptor_radius / 10.0;      // src:351
  const double point_density_radius = P.descriptor_radius / 5.0;  // src:352
  ShapeContextTables T;
  sc3d_init_compute(search_radius, min_radius, T);
  const float R2f = radius_sq_float(search_radius);
  const float rho2f = radius_sq_float(point_density_radius);

  Searcher tree;  // pcl::search::KdTree default-constructed: sorted results (src:335,348)
  tree.set_input(cloud.data(), (int)cloud.size(), use_tree, /*sorted=*/true);

  // boost::uniform_01<boost::mt19937> seeded 12345u, fresh per ShapeContext3DEstimation (src:343)
  std::mt19937 rng_alg(12345u);
  auto rnd = [&]() -> double { return (double)rng_alg() * (1.0 / 4294967296.0); };

  std::vector<Hit> nn, nn2;
  std::vector<int> rho_cache;  // density depends only on the surface point; cached in brute mode
  if (!use_tree) rho_cache.assign(cloud.size(), -1);
  // boundary report: the density search of a surface point is counted once, however many keypoints reach it
  std::vector<char> rho_counted;
  if (bc) rho_counted.assign(cloud.size(), 0);

  for (size_t kp = 0; kp < keypoints.size(); kp++) {
    float* d = desc + kp * FE_DESC_LEN;
    if (dbg) { dbg[kp].edge_margin = std::numeric_limits<float>::infinity(); dbg[kp].n_neighbors = 0; }
    const P4& in = keypoints[kp];
    if (!std::isfinite(in.x) || !std::isfinite(in.y) || !std::isfinite(in.z)) {
      for (int i = 0; i < FE_DESC_LEN; i++) d[i] = std::numeric_limits<float>::quiet_NaN();
      continue;
    }
    if (bc) { tree.bnd_eps = bc->eps; tree.bnd_hits = 0; }
    const size_t neighb_cnt = (size_t)tree.radius(in, R2f, nn);
    if (bc) { bc->n[2] += tree.bnd_hits; tree.bnd_eps = 0.0; }
    if (neighb_cnt == 0) {
      for (int i = 0; i < FE_DESC_LEN; i++) d[i] = std::numeric_limits<float>::quiet_NaN();
      continue;
    }
    if (dbg) dbg[kp].n_neighbors = (int)neighb_cnt;
    for (int i = 0; i < FE_DESC_LEN; i++) d[i] = 0.0f;
    const V3 origin = v3(in.x, in.y, in.z);
    V3 normal = normal_const;  // normals[minIndex]: all equal
    V3 x_axis;
    x_axis.x = static_cast<float>(rnd());
    x_axis.y = static_cast<float>(rnd());
    x_axis.z = static_cast<float>(rnd());
    if (!pcl_utils_equal(normal.z, 0.0f))
      x_axis.z = -(normal.x * x_axis.x + normal.y * x_axis.y) / normal.z;
    else if (!pcl_utils_equal(normal.y, 0.0f))
      x_axis.y = -(normal.x * x_axis.x + normal.z * x_axis.z) / normal.y;
    else if (!pcl_utils_equal(normal.x, 0.0f))
      x_axis.x = -(normal.y * x_axis.y + normal.z * x_axis.z) / normal.x;
    eigen32_normalize(x_axis);

    for (size_t ne = 0; ne < neighb_cnt; ne++) {
      if (bc && !rho_counted[nn[ne].idx]) {  // boundary report only: a search of its own, once per surface point
        rho_counted[nn[ne].idx] = 1;
        tree.bnd_eps = bc->eps; tree.bnd_hits = 0;
        tree.radius(cloud[nn[ne].idx], rho2f, nn2);
        bc->n[3] += tree.bnd_hits;
        tree.bnd_eps = 0.0;
      }
      if (pcl_utils_equal(nn[ne].d2, 0.0f)) continue;
      const P4& nbp = cloud[nn[ne].idx];
      const V3 neighbour = v3(nbp.x, nbp.y, nbp.z);
      float r = sqrtf(nn[ne].d2);
      // pcl::geometry::project(neighbour, origin, normal, proj); proj -= origin;
      V3 po = sub(neighbour, origin);
      float lambda = eigen_dot(normal, po);
      V3 proj = v3(neighbour.x - lambda * normal.x, neighbour.y - lambda * normal.y, neighbour.z - lambda * normal.z);
      proj = sub(proj, origin);
      eigen32_normalize(proj);
      V3 cross = eigen_cross(x_axis, proj);
      float phi = pcl_rad2deg(atan2f(eigen_norm(cross), eigen_dot(x_axis, proj)));
      phi = eigen_dot(cross, normal) < 0.f ? (360.0f - phi) : phi;
      V3 no = sub(neighbour, origin);
      eigen32_normalize(no);
      float theta = eigen_dot(normal, no);
      theta = pcl_rad2deg(acosf(std::min(1.0f, std::max(-1.0f, theta))));

      size_t j = 0, k = 0, l = 0;
      for (size_t rad = 1; rad < 15 + 1; rad++) if (r <= T.radii[rad]) { j = rad - 1; break; }
      for (size_t ang = 1; ang < 11 + 1; ang++) if (theta <= T.theta[ang]) { k = ang - 1; break; }
      for (size_t ang = 1; ang < 12 + 1; ang++) if (phi <= T.phi[ang]) { l = ang - 1; break; }

      if (dbg) {
        float m = dbg[kp].edge_margin;
        for (int a = 1; a < 12; a++) { float s = std::max(fabsf(theta), 1e-3f); m = std::min(m, fabsf(theta - T.theta[a]) / s); }
        for (int a = 1; a < 13; a++) { float s = std::max(fabsf(phi), 1e-3f); m = std::min(m, fabsf(phi - T.phi[a]) / s); }
        if (!(m == m)) m = 0.0f;
        dbg[kp].edge_margin = m;
      }

      int point_density;
      if (!use_tree && rho_cache[nn[ne].idx] >= 0) point_density = rho_cache[nn[ne].idx];
      else {
        point_density = tree.radius(nbp, rho2f, nn2);
        if (!use_tree) rho_cache[nn[ne].idx] = point_density;
      }
      if (point_density == 0) continue;
      float w = (1.0f / static_cast<float>(point_density)) * T.lut[(l * 11 * 15) + (k * 15) + j];
      d[(l * 11 * 15) + (k * 15) + j] += w;
    }
  }
}

// ------------------------------------------------------------------------------------------
// cloudCallback, src:83-117.
// ------------------------------------------------------------------------------------------
struct ScanResult {
  std::vector<P4> cloud_full, cloud, keypoints, keypoint_cloud;
  std::vector<float> descriptors;
  std::vector<DescDebug> dbg;
  BoundaryCounter bnd;
};

void process_scan(const fe_params_t& P, const P4* pts, int64_t n, double roll, double pitch,
                  bool use_tree, bool want_dbg, ScanResult& R, double boundary_eps = 0.0) {
  R.bnd = BoundaryCounter();
  R.bnd.eps = boundary_eps;
  BoundaryCounter* bc = boundary_eps > 0.0 ? &R.bnd : nullptr;
  R.cloud_full.assign(pts, pts + n);
  get_elevation_angles(R.cloud_full.data(), n);      // src:87
  rotate_cloud(R.cloud_full.data(), n, roll, pitch); // src:92
  R.cloud = R.cloud_full;                            // src:98
  filter_cloud(P, R.cloud);                          // src:99
  R.keypoints.clear(); R.keypoint_cloud.clear();
  estimate_keypoints(P, R.cloud, use_tree, R.keypoints, R.keypoint_cloud, nullptr, bc);  // src:107
  R.descriptors.clear(); R.dbg.clear();
  if (P.estimate_descriptors) {  // src:112
    R.descriptors.assign(R.keypoints.size() * (size_t)FE_DESC_LEN, 0.0f);
    if (want_dbg) R.dbg.resize(R.keypoints.size());
    estimate_descriptors(P, R.cloud_full, R.keypoints, use_tree, R.descriptors.data(),
                         want_dbg ? R.dbg.data() : nullptr, bc);  // src:115
  }
}

template <class T>
int copy_out(const std::vector<T>& v, T* out, int64_t cap, int64_t* n_out) {
  if (n_out) *n_out = (int64_t)v.size();
  if (!out) return FE_OK;
  if ((int64_t)v.size() > cap) return FE_ERR_CAPACITY;
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(T));
  return FE_OK;
}

}  // namespace

// ==========================================================================================
// C entry points (feo_*): the oracle's mirror of include/fe_b200.h.  mode: 0 = brute force,
// 1 = KD-tree (leaf 15).
// ==========================================================================================
extern "C" {

const char* feo_version(void) { return "fe_oracle 0.1 (CPU restatement; parity unpinned)"; }

void feo_params_node_default(fe_params_t* p) {  // src:9-34
  p->x_min = 0.0; p->x_max = 75.0; p->y_min = -30.0; p->y_max = 30.0; p->z_min = -1.5; p->z_max = 5.0;
  p->cluster_tolerance = 0.65; p->cluster_min_count = 5; p->cluster_max_count = 50;
  p->cluster_radius_threshold = 0.15; p->number_detection_channels = 1;
  p->estimate_descriptors = 1; p->descriptor_radius = 2.5;
}

void feo_params_launch_playback(fe_params_t* p) {  // launch/keypoint_playback.launch:17-33
  feo_params_node_default(p);
  p->x_min = 0.0; p->x_max = 100.0; p->y_min = -50.0; p->y_max = 50.0; p->z_min = -1.5; p->z_max = 4.0;
  p->cluster_tolerance = 1.0; p->cluster_min_count = 1; p->cluster_max_count = 1000;
  p->cluster_radius_threshold = 0.2; p->number_detection_channels = 2; p->descriptor_radius = 2.5;
}

int feo_get_elevation_angles(fe_point_t* cloud, int64_t n) { get_elevation_angles(cloud, n); return FE_OK; }

int feo_rotation_matrix(double roll, double pitch, float m[9]) { rotation_matrix(roll, pitch, m); return FE_OK; }

int feo_rotate_cloud(fe_point_t* cloud, int64_t n, double roll, double pitch) {
  rotate_cloud(cloud, n, roll, pitch); return FE_OK;
}

int feo_filter_cloud(const fe_params_t* P, const fe_point_t* in, int64_t n, fe_point_t* out,
                     int64_t cap, int64_t* n_out) {
  std::vector<P4> c(in, in + n);
  filter_cloud(*P, c);
  return copy_out(c, out, cap, n_out);
}

// ring selection of estimateKeypoints (src:200-202) for ring i in 0..15
int feo_select_ring(const fe_point_t* cloud, int64_t n, int ring, fe_point_t* out, int64_t cap, int64_t* n_out) {
  std::vector<P4> c(cloud, cloud + n), ch;
  double c0 = (ring - 7) * 2 - 1;
  pass_through(c, FI, c0 - 1.0, c0 + 1.0, ch);
  return copy_out(ch, out, cap, n_out);
}

int feo_extract_clusters(const fe_point_t* cloud, int64_t n, double tolerance, int32_t min_size,
                         int32_t max_size, int32_t mode, int32_t* cluster_offsets,
                         int32_t cap_clusters, int32_t* indices, int64_t cap_indices,
                         int32_t* n_clusters) {
  std::vector<P4> c(cloud, cloud + n);
  Clusters cl;
  euclidean_cluster_extract(c, tolerance, min_size, max_size, mode == 1, cl);
  *n_clusters = (int32_t)cl.size();
  if ((int32_t)cl.size() > cap_clusters) return FE_ERR_CAPACITY;
  int64_t t = 0;
  cluster_offsets[0] = 0;
  for (size_t i = 0; i < cl.size(); i++) {
    if (t + (int64_t)cl[i].size() > cap_indices) return FE_ERR_CAPACITY;
    for (size_t j = 0; j < cl[i].size(); j++) indices[t++] = cl[i][j];
    cluster_offsets[i + 1] = (int32_t)t;
  }
  return FE_OK;
}

int feo_get_cylinder_segments(const fe_params_t* P, const fe_point_t* ring_cloud, int64_t n, int32_t mode,
                              fe_point_t* centroids, int64_t cap_centroids, int64_t* n_centroids,
                              fe_point_t* cluster_cloud, int64_t cap_cloud, int64_t* n_cloud) {
  std::vector<P4> c(ring_cloud, ring_cloud + n), kp, kc;
  get_cylinder_segments(*P, c, mode == 1, kp, kc);
  int s = copy_out(kp, centroids, cap_centroids, n_centroids);
  if (s) return s;
  return copy_out(kc, cluster_cloud, cap_cloud, n_cloud);
}

int feo_estimate_keypoints(const fe_params_t* P, const fe_point_t* cloud, int64_t n, int32_t mode,
                           fe_point_t* keypoints, int64_t cap_keypoints, int64_t* n_keypoints,
                           fe_point_t* keypoint_cloud, int64_t cap_cloud, int64_t* n_cloud,
                           fe_point_t* keypoints_full, int64_t cap_full, int64_t* n_full) {
  std::vector<P4> c(cloud, cloud + n), kp, kc, kf;
  estimate_keypoints(*P, c, mode == 1, kp, kc, &kf);
  int s = copy_out(kp, keypoints, cap_keypoints, n_keypoints);
  if (s) return s;
  s = copy_out(kc, keypoint_cloud, cap_cloud, n_cloud);
  if (s) return s;
  return copy_out(kf, keypoints_full, cap_full, n_full);
}

// edge_margin (nullable): per keypoint, min relative distance of any neighbour's theta/phi to a
// bin edge — the only place where libm (acosf/atan2f) differences can move a contribution.
int feo_estimate_descriptors(const fe_params_t* P, const fe_point_t* cloud_full, int64_t n,
                             const fe_point_t* keypoints, int64_t k, int32_t mode,
                             float* descriptors, float* edge_margin, int32_t* n_neighbors) {
  std::vector<P4> c(cloud_full, cloud_full + n), kp(keypoints, keypoints + k);
  std::vector<DescDebug> dbg(k);
  estimate_descriptors(*P, c, kp, mode == 1, descriptors, dbg.data());
  for (int64_t i = 0; i < k; i++) {
    if (edge_margin) edge_margin[i] = dbg[i].edge_margin;
    if (n_neighbors) n_neighbors[i] = dbg[i].n_neighbors;
  }
  return FE_OK;
}

int feo_sc3d_tables(double search_radius, float radii[16], float theta[12], float phi[13], float* lut) {
  ShapeContextTables T;
  sc3d_init_compute(search_radius, search_radius / 10.0, T);
  memcpy(radii, T.radii, sizeof(T.radii));
  memcpy(theta, T.theta, sizeof(T.theta));
  memcpy(phi, T.phi, sizeof(T.phi));
  if (lut) memcpy(lut, T.lut, sizeof(T.lut));
  return FE_OK;
}

float feo_radius_sq_float(double tol_as_passed, int32_t narrow_first) {
  // narrow_first=1: EuclideanClusterExtraction path (tolerance narrowed to float first)
  if (narrow_first) return radius_sq_float((double)(float)tol_as_passed);
  return radius_sq_float(tol_as_passed);
}

void feo_mt19937_draws(uint32_t seed, int32_t n, uint32_t* raw, float* as_float) {
  std::mt19937 g(seed);
  for (int i = 0; i < n; i++) {
    uint32_t v = (uint32_t)g();
    if (raw) raw[i] = v;
    if (as_float) as_float[i] = static_cast<float>((double)v * (1.0 / 4294967296.0));
  }
}

// The order std::sort(rbegin, rend, size<) leaves `n` clusters of the given sizes in:
// order_out[p] = original position of the cluster that ends at position p.
void feo_std_sort_cluster_order(const int32_t* sizes, int32_t n, int32_t* order_out) {
  struct C { int size, id; };
  std::vector<C> v(n);
  for (int i = 0; i < n; i++) { v[i].size = sizes[i]; v[i].id = i; }
  std::sort(v.rbegin(), v.rend(), [](const C& a, const C& b) { return a.size < b.size; });
  for (int i = 0; i < n; i++) order_out[i] = v[i].id;
}

// imuCallback, src:57-70: tf::quaternionMsgToTF + tf::Matrix3x3(quat).getRPY(tmproll, pitch, yaw)
// [tf/LinearMath/Matrix3x3.h, not vendored: setRotation + getEulerYPR solution 1, tfScalar = double].
int feo_imu_to_roll_pitch(const double q[4], int32_t levelCloud, double* roll_out, double* pitch_out) {
  double roll, pitch;
  if (levelCloud) {
    const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    double d = qx * qx + qy * qy + qz * qz + qw * qw;  // Quaternion::length2
    double s = 2.0 / d;
    double xs = qx * s, ys = qy * s, zs = qz * s;
    double wx = qw * xs, wy = qw * ys;
    double xx = qx * xs, xz = qx * zs;
    double yy = qy * ys, yz = qy * zs;
    double m_el2_x = xz - wy, m_el2_y = yz + wx, m_el2_z = 1.0 - (xx + yy);
    double tmproll;
    if (fabs(m_el2_x) >= 1) {  // gimbal lock
      double delta = atan2(m_el2_y, m_el2_z);
      if (m_el2_x < 0) { pitch = M_PI / 2.0; tmproll = delta; }
      else { pitch = -M_PI / 2.0; tmproll = delta; }
    } else {
      pitch = -asin(m_el2_x);
      tmproll = atan2(m_el2_y / cos(pitch), m_el2_z / cos(pitch));
    }
    roll = tmproll - M_PI;  // src:65
  } else {
    roll = 0.0;   // src:67-68
    pitch = 0.0;
  }
  *roll_out = roll; *pitch_out = pitch;
  return FE_OK;
}

// One scan through the whole path (cloudCallback, src:83-117).  Any output pointer may be NULL.
int feo_process_scan(const fe_params_t* P, const fe_point_t* points, int64_t n, double roll,
                     double pitch, int32_t mode,
                     fe_point_t* keypoints, int64_t cap_kp, int64_t* n_kp,
                     float* descriptors, float* edge_margin,
                     fe_point_t* keypoint_cloud, int64_t cap_kc, int64_t* n_kc,
                     fe_point_t* cloud, int64_t cap_cloud, int64_t* n_cloud,
                     fe_point_t* cloud_full) {
  ScanResult R;
  process_scan(*P, points, n, roll, pitch, mode == 1, edge_margin != nullptr, R);
  int s = copy_out(R.keypoints, keypoints, cap_kp, n_kp);
  if (s) return s;
  if (descriptors && !R.descriptors.empty())
    memcpy(descriptors, R.descriptors.data(), R.descriptors.size() * sizeof(float));
  if (edge_margin) for (size_t i = 0; i < R.dbg.size(); i++) edge_margin[i] = R.dbg[i].edge_margin;
  s = copy_out(R.keypoint_cloud, keypoint_cloud, cap_kc, n_kc);
  if (s) return s;
  s = copy_out(R.cloud, cloud, cap_cloud, n_cloud);
  if (s) return s;
  if (cloud_full && n > 0) memcpy(cloud_full, R.cloud_full.data(), (size_t)n * sizeof(P4));
  return FE_OK;
}

// A batch of scans, scan-parallel over n_threads host threads (1 = how the reference runs:
// one scan at a time in ros::spin, src:386).  keypoint_offsets has n_scans+1 entries.
// descriptors (nullable) must hold cap_kp * 1980 floats.  Returns FE_ERR_CAPACITY on overflow
// (n_kp_total is still set to the required total).
static int process_batch_impl(const fe_params_t* P, const fe_point_t* points, const int64_t* scan_offsets,
                      const double* roll_pitch, int32_t n_scans, int32_t mode, int32_t n_threads,
                      int64_t* keypoint_offsets, fe_point_t* keypoints, float* descriptors,
                      float* edge_margin, int64_t cap_kp, int64_t* n_kp_total,
                      double boundary_eps, int64_t* boundary_counts) {
  std::vector<ScanResult> res(n_scans);
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;) {
      int s = next.fetch_add(1);
      if (s >= n_scans) break;
      process_scan(*P, points + scan_offsets[s], scan_offsets[s + 1] - scan_offsets[s],
                   roll_pitch[2 * s], roll_pitch[2 * s + 1], mode == 1, edge_margin != nullptr, res[s],
                   boundary_counts ? boundary_eps : 0.0);
      res[s].cloud_full.clear(); res[s].cloud_full.shrink_to_fit();
      res[s].cloud.clear(); res[s].cloud.shrink_to_fit();
      res[s].keypoint_cloud.clear();
    }
  };
  if (n_threads == 1) worker();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  int64_t total = 0;
  if (keypoint_offsets) keypoint_offsets[0] = 0;
  for (int s = 0; s < n_scans; s++) {
    total += (int64_t)res[s].keypoints.size();
    if (keypoint_offsets) keypoint_offsets[s + 1] = total;
  }
  if (n_kp_total) *n_kp_total = total;
  if (boundary_counts)
    for (int s = 0; s < n_scans; s++)
      for (int k = 0; k < 4; k++) boundary_counts[4 * s + k] = res[s].bnd.n[k];
  if (!keypoints) return FE_OK;
  if (total > cap_kp) return FE_ERR_CAPACITY;
  int64_t t = 0;
  for (int s = 0; s < n_scans; s++) {
    size_t k = res[s].keypoints.size();
    if (!k) continue;
    memcpy(keypoints + t, res[s].keypoints.data(), k * sizeof(P4));
    if (descriptors && !res[s].descriptors.empty())
      memcpy(descriptors + t * FE_DESC_LEN, res[s].descriptors.data(), k * FE_DESC_LEN * sizeof(float));
    if (edge_margin) for (size_t i = 0; i < res[s].dbg.size(); i++) edge_margin[t + i] = res[s].dbg[i].edge_margin;
    t += (int64_t)k;
  }
  return FE_OK;
}

int feo_process_batch(const fe_params_t* P, const fe_point_t* points, const int64_t* scan_offsets,
                      const double* roll_pitch, int32_t n_scans, int32_t mode, int32_t n_threads,
                      int64_t* keypoint_offsets, fe_point_t* keypoints, float* descriptors,
                      float* edge_margin, int64_t cap_kp, int64_t* n_kp_total) {
  return process_batch_impl(P, points, scan_offsets, roll_pitch, n_scans, mode, n_threads, keypoint_offsets, keypoints,
                            descriptors, edge_margin, cap_kp, n_kp_total, 0.0, nullptr);
}

// The same batch with the tolerance-boundary report: boundary_counts[4*s + k] = pairs of scan s within
// eps_m of the radius of predicate k (0 ring clustering, 1 cross-ring merge, 2 3DSC support radius,
// 3 3DSC point-density radius).
int feo_process_batch_boundary(const fe_params_t* P, const fe_point_t* points, const int64_t* scan_offsets,
                               const double* roll_pitch, int32_t n_scans, int32_t mode, int32_t n_threads,
                               double eps_m, int64_t* boundary_counts) {
  return process_batch_impl(P, points, scan_offsets, roll_pitch, n_scans, mode, n_threads, nullptr, nullptr, nullptr,
                            nullptr, 0, nullptr, eps_m, boundary_counts);
}

// The host libm's float routines, element-wise: op 0 = atan2f(a, b), 1 = acosf(a), 2 = atanf(a).
// What tests/ compare the device's restatement (csrc/glibc_f32.h) against.
void feo_libm_f32(int32_t op, const float* a, const float* b, float* out, int64_t n) {
  for (int64_t i = 0; i < n; i++) {
    volatile float x = a[i];
    out[i] = op == 0 ? atan2f(x, b[i]) : op == 1 ? acosf(x) : atanf(x);
  }
}

}  // extern "C"

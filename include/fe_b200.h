/*
 * fe_b200.h — C-ABI of the B200-native per-scan keypoint pipeline.
 *
 * Drop-in boundary for the span of `cloudCallback` between
 * src/feature_extraction_node.cpp:83 (after PointCloud2 -> pcl conversion) and :117
 * (before publishing) of GAVLab/feature_extraction.  ROS / PointCloud2 I/O stays on the
 * host and is not part of this library.
 *
 * Everything behind these entry points runs as hand-written sm_100a CUDA kernels.  There is
 * NO CPU fallback: fe_create() fails when no CUDA device is usable.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary; every call returns an fe_status.
 *   - a context (fe_ctx_t) is bound to one GPU; it is NOT thread-safe; distinct contexts are
 *     independent (one per GPU for scan-parallel sharding).  The reference is single-threaded
 *     and serial per scan (ros::spin, src:386).
 *   - empty in -> empty out is not an error (src:209-210, 234-235, 263-264, 278-279, 331-332).
 *   - capacity overflow is FE_ERR_CAPACITY, never silent truncation.
 */
#ifndef FE_B200_H_
#define FE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FE_DESC_LEN 1980          /* pcl::ShapeContext1980: 12 az x 11 el x 15 rad          */
#define FE_RECORD_FLOATS 1996     /* pcl::PointDescriptor, feature_extraction_node.h:35-42:  */
                                  /* x,y,z,pad | intensity | descriptor[1980] | rf[9] -> 7976 B, 16-aligned 7984 B */
#define FE_NUM_RINGS 16           /* hard-coded channel loop, src:195                        */

typedef enum fe_status {
  FE_OK = 0,
  FE_ERR_INVALID = 1,       /* bad argument                                              */
  FE_ERR_CUDA = 2,          /* CUDA runtime error (see fe_last_error)                    */
  FE_ERR_CAPACITY = 3,      /* an output or workspace capacity would be exceeded         */
  FE_ERR_NO_DEVICE = 4,     /* no usable CUDA device: there is no CPU fallback           */
  FE_ERR_UNSUPPORTED = 5
} fe_status;

/* pcl::PointXYZI as the reference uses it (feature_extraction_node.h:66) minus PCL's padding:
 * 16 bytes, `intensity` carries the elevation angle in degrees after getElevationAngles. */
typedef struct fe_point {
  float x, y, z, intensity;
} fe_point_t;

/* The ROS parameters read in the constructor, src:9-34, with the member types of
 * feature_extraction_node.h:115-127.  `cloud_leveling` is applied by the caller (pass roll =
 * pitch = 0, as imuCallback does at src:66-69). */
typedef struct fe_params {
  double x_min, x_max;                /* src:14-15 */
  double y_min, y_max;                /* src:16-17 */
  double z_min, z_max;                /* src:18-19 */
  double cluster_tolerance;           /* src:24 */
  int32_t cluster_min_count;          /* src:25 */
  int32_t cluster_max_count;          /* src:26 */
  double cluster_radius_threshold;    /* src:27 */
  int32_t number_detection_channels;  /* src:28 */
  int32_t estimate_descriptors;       /* src:33 (bool) */
  double descriptor_radius;           /* src:34 */
} fe_params_t;

/* Workspace capacities of a context.  0 = library default. */
typedef struct fe_limits {
  int64_t max_points_per_call;     /* points staged on the device per sub-batch            */
  int32_t max_scans_per_call;      /* scans per sub-batch                                  */
  int64_t max_keypoints_per_call;  /* keypoints (and descriptors) per sub-batch            */
  int64_t max_ring_clusters_per_call; /* ring-level centroids (keypoints_full, src:205)    */
} fe_limits_t;

typedef struct fe_ctx fe_ctx_t;

/* ---- parameter presets ------------------------------------------------------------------ */
void fe_params_node_default(fe_params_t* p);     /* constructor defaults, src:9-34            */
void fe_params_launch_playback(fe_params_t* p);  /* launch/keypoint_playback.launch:17-33     */

/* ---- lifecycle --------------------------------------------------------------------------- */
const char* fe_version(void);
int fe_create(int device, const fe_params_t* params, const fe_limits_t* limits, fe_ctx_t** out);
int fe_set_params(fe_ctx_t* ctx, const fe_params_t* params);
void fe_destroy(fe_ctx_t* ctx);
const char* fe_last_error(const fe_ctx_t* ctx);
int fe_device_count(void);

/* Pinned host memory helpers (optional; any host pointer is accepted by the batch call, pinned
 * memory lets the H2D/D2H copies overlap the kernels). */
void* fe_host_alloc(int64_t bytes);
void fe_host_free(void* p);

/* ---- the fused path: cloudCallback, src:83-117 ------------------------------------------- */

typedef struct fe_batch_result {
  int32_t n_scans;
  int64_t n_keypoints;
  const int64_t* keypoint_offsets; /* host, n_scans+1 (CSR by scan)                         */
  const fe_point_t* keypoints;     /* n_keypoints x {x,y,z,el_deg}  (~keypoints, src:131)   */
  const float* descriptors;        /* n_keypoints x 1980 (NULL if estimate_descriptors==0)  */
  int32_t on_device;               /* 1: keypoints/descriptors are device pointers          */
  int64_t gpu_launches;            /* kernels launched by this call                         */
} fe_batch_result_t;

/* points/scan_offsets/roll_pitch are HOST buffers.  Scan s owns points
 * [scan_offsets[s], scan_offsets[s+1]); roll_pitch[2s], roll_pitch[2s+1] are the IMU roll and
 * pitch in radians as imuCallback leaves them (src:63-65).  Results are context-owned host
 * memory, valid until the next call on this context. */
int fe_process_batch(fe_ctx_t* ctx, const fe_point_t* points, const int64_t* scan_offsets,
                     const double* roll_pitch, int32_t n_scans, fe_batch_result_t* out);

/* Layout of the caller's point records (sensor_msgs/PointCloud2 point_step and field offsets, or any
 * array of structs): x, y, z are little-endian 32-bit floats at byte offsets x_off, y_off, z_off of
 * records `stride` bytes apart.  The input intensity is never read: getElevationAngles overwrites it
 * (src:154).  Examples: {16,0,4,8} fe_point_t; {12,0,4,8} packed xyz; {32,0,4,8} pcl::PointXYZI as
 * it sits in a pcl::PointCloud; {22,0,4,8} / {32,0,4,8} Velodyne PointCloud2 data (src:79-81). */
typedef struct fe_point_layout {
  int32_t stride, x_off, y_off, z_off;
} fe_point_layout_t;

/* fe_process_batch for records of any layout: `points` is the HOST byte buffer, scan_offsets count
 * records.  The records are copied to the device as they are and decoded by the first kernel. */
int fe_process_batch_layout(fe_ctx_t* ctx, const void* points, const fe_point_layout_t* layout,
                            const int64_t* scan_offsets, const double* roll_pitch, int32_t n_scans,
                            fe_batch_result_t* out);

/* imuCallback, src:57-70: orientation quaternion {x,y,z,w} -> the roll/pitch the node keeps
 * (tf::Matrix3x3(quat).getRPY, roll = tmproll - pi; both 0 when cloud_leveling is false). */
int fe_imu_to_roll_pitch(const double quat_xyzw[4], int32_t cloud_leveling, double* roll, double* pitch);

/* Same, but `d_points` is already resident in device memory and the results stay on the
 * device (keypoint_offsets is still host memory).  Must fit one sub-batch.
 * Calls of a handful of scans (<= 16; also through fe_process_batch) are replayed from a CUDA graph
 * captured the second time a shape is seen — here the shape includes `d_points`, so a caller that
 * keeps refilling the same device buffer gets the replay.  Results do not depend on the route. */
int fe_process_batch_device(fe_ctx_t* ctx, const fe_point_t* d_points,
                            const int64_t* scan_offsets, const double* roll_pitch,
                            int32_t n_scans, fe_batch_result_t* out);

/* ---- scan-parallel sharding over several GPUs of one box, in one process -------- */
/* Scans are independent (src:72-145 reads no cross-scan state): GPU g of G takes the contiguous scan
 * range [g*B/G, (g+1)*B/G), one host thread and one context per GPU, and the per-GPU results are
 * concatenated on the host in scan order.  No collective is involved. */
typedef struct fe_multi fe_multi_t;
int fe_multi_create(const int32_t* devices, int32_t n_devices, const fe_params_t* params,
                    const fe_limits_t* limits, fe_multi_t** out);
int fe_multi_process_batch(fe_multi_t* m, const fe_point_t* points, const int64_t* scan_offsets,
                           const double* roll_pitch, int32_t n_scans, fe_batch_result_t* out);
void fe_multi_destroy(fe_multi_t* m);
const char* fe_multi_last_error(const fe_multi_t* m);

/* Copy `bytes` of a device-resident result (fe_process_batch_device) to host memory. */
int fe_download(fe_ctx_t* ctx, void* host_dst, const void* device_src, int64_t bytes);

/* Optional extra outputs of the last fe_process_batch() / fe_process_batch_layout() (any number of
 * sub-batches): ~keypoint_cloud (src:133-135, the points of every gated ring cluster in the order
 * src:323 appends them) and ~cloud (src:137-139, the cropped cloud), CSR by scan, gathered on the device
 * and copied into context-owned pinned memory (valid until the next call).  Enabled with
 * fe_enable_cloud_outputs(ctx, 1) before the call; fe_multi_* concatenates the shards' outputs. */
int fe_enable_cloud_outputs(fe_ctx_t* ctx, int32_t enable);
int fe_get_cloud_outputs(fe_ctx_t* ctx, const int64_t** cloud_offsets, const fe_point_t** cloud,
                         const int64_t** kpcloud_offsets, const fe_point_t** keypoint_cloud);
int fe_multi_enable_cloud_outputs(fe_multi_t* m, int32_t enable);
int fe_multi_get_cloud_outputs(fe_multi_t* m, const int64_t** cloud_offsets, const fe_point_t** cloud,
                               const int64_t** kpcloud_offsets, const fe_point_t** keypoint_cloud);

/* Record output — pcl::concatenateFields(*keypoints, *descriptors, *pt_descriptors) at src:119 done
 * on the device: when enabled, `descriptors` of every following result (fe_process_batch,
 * fe_process_batch_layout, fe_process_batch_device, fe_multi_process_batch) points to n_keypoints
 * records of FE_RECORD_FLOATS floats in the layout of pcl::PointDescriptor
 * (feature_extraction_node.h:35-53): x, y, z, 0 | intensity | descriptor[1980] | rf[9] = 0 | padding —
 * what fe_pack_point_descriptors() would build on the host, ready to be published as ~features.
 * `keypoints` is filled as before.  With estimate_descriptors == 0 there are no descriptors and hence no
 * records (the reference publishes ~features only inside `if (descriptorEstimation)`, src:112-124): the
 * setting is kept but has no effect. */
int fe_enable_record_output(fe_ctx_t* ctx, int32_t enable);
int fe_multi_enable_record_output(fe_multi_t* m, int32_t enable);

/* Which float libm the 3DSC angles follow.  pcl::ShapeContext3DEstimation::computePoint (3dsc.hpp,
 * called at src:353) bins a neighbour by rad2deg(atan2(float, float)) and rad2deg(acosf(float)), i.e. by
 * the last bit of the host libm's atan2f / acosf.  FE_LIBM_FDLIBM (default): the fdlibm float routines
 * of glibc <= 2.40 restated on the device operation by operation (csrc/glibc_f32.h) — bit-identical to
 * the reference's x86-64 build and to this image's glibc 2.39.  FE_LIBM_CORRECTLY_ROUNDED: what
 * glibc >= 2.41 (CORE-MATH) returns. */
#define FE_LIBM_FDLIBM 0
#define FE_LIBM_CORRECTLY_ROUNDED 1
int fe_set_angle_libm(fe_ctx_t* ctx, int32_t mode);

/* Tolerance-boundary report (BASELINE.json north_star: "points lying within 1e-6 m of a tolerance
 * boundary reported separately").  Every radius predicate of the path is d2 < r2f evaluated in float
 * (FLANN L2_Simple; r2f = (float)(radius*radius)); a point pair is ON THE BOUNDARY of a predicate when
 * |sqrt((double)d2) - sqrt((double)r2f)| < eps_m.  After fe_enable_boundary_report(ctx, eps_m > 0) every
 * batch call also counts those pairs per scan — audit kernels next to the pipeline, slower — and
 * fe_get_boundary_report returns n_scans x 4 counts of the last call:
 *   [0] ring clustering, src:269-276: unordered pairs of cropped points of one ring (a point on a ring
 *       window's end value belongs to two rings, src:200-202, and counts in both)
 *   [1] cross-ring merge, src:222-229: unordered pairs of ring centroids with the pseudo z of src:217
 *   [2] 3DSC support radius, src:350: (keypoint, surface point) pairs
 *   [3] 3DSC point-density radius, src:352: (neighbour, surface point) pairs, the neighbour being every
 *       surface point inside some keypoint's support sphere, counted once
 * A scan with all four counts 0 has no decision within eps_m of flipping; the CPU oracle counts the same
 * pairs (feo_process_batch_boundary) and tools/parity_campaign.py prints both.  eps_m <= 0 disables. */
int fe_enable_boundary_report(fe_ctx_t* ctx, double eps_m);
int fe_get_boundary_report(fe_ctx_t* ctx, const int64_t** counts, int32_t* n_scans);

/* CUDA-event stopwatch on the context's stream (the stream every kernel of
 * fe_process_batch_device is launched on): begin, run any number of calls, end -> elapsed ms. */
int fe_timer_begin(fe_ctx_t* ctx);
int fe_timer_end(fe_ctx_t* ctx, float* elapsed_ms);

/* Work counters of the last fe_process_batch_device call (for roofline arithmetic):
 * out[0..9] = points, surface points kept, cropped points, ring clusters (keypoints_full),
 * keypoints, sum of 3DSC neighbours, scans deferred to the large K2, to the large K3, to the
 * global-memory K4a, keypoints whose descriptor was summed out of PCL's order (a single bin with
 * more than 8192 contributions; larger neighbourhoods are otherwise handled exactly, in bin groups). */
int fe_get_batch_stats(fe_ctx_t* ctx, int64_t out[10]);

/* Per-kernel CUDA-event times (ms) of the last device call; names are static strings.  Only collected
 * after fe_enable_stage_timing(ctx, 1) (an event pair around every stage; CUDA-graph replay is
 * bypassed while it is on). */
int fe_enable_stage_timing(fe_ctx_t* ctx, int32_t enable);
int fe_get_stage_times(fe_ctx_t* ctx, int32_t cap, const char** names, float* ms, int32_t* n);

/* ---- one entry point per reference function (host buffers, synchronous) ------------------ */

/* getElevationAngles, src:147-156: overwrites intensity with the elevation angle (deg). */
int fe_get_elevation_angles(fe_ctx_t* ctx, fe_point_t* cloud, int64_t n);

/* rotateCloud, src:159-167 (pcl::transformPointCloud with
 * AngleAxisf(pitch,Y)*AngleAxisf(roll,X)); in place, intensity carried through. */
int fe_rotate_cloud(fe_ctx_t* ctx, fe_point_t* cloud, int64_t n, double roll, double pitch);

/* The 3x3 float rotation rotateCloud applies (row-major), for inspection. */
int fe_rotation_matrix(double roll, double pitch, float m[9]);

/* filterCloud, src:169-183 (three pcl::PassThrough passes): stable crop. */
int fe_filter_cloud(fe_ctx_t* ctx, const fe_point_t* in, int64_t n, fe_point_t* out,
                    int64_t cap, int64_t* n_out);

/* pcl::EuclideanClusterExtraction::extract as called at src:222-229 and src:269-276.
 * cluster c owns indices[cluster_offsets[c] .. cluster_offsets[c+1]) (ascending), clusters in
 * PCL's output order. */
int fe_extract_clusters(fe_ctx_t* ctx, const fe_point_t* cloud, int64_t n, double tolerance,
                        int32_t min_size, int32_t max_size, int32_t* cluster_offsets,
                        int32_t cap_clusters, int32_t* indices, int64_t cap_indices,
                        int32_t* n_clusters);

/* getCylinderSegments, src:261-327: one ring's cloud -> gated centroids + their points. */
int fe_get_cylinder_segments(fe_ctx_t* ctx, const fe_point_t* ring_cloud, int64_t n,
                             fe_point_t* centroids, int64_t cap_centroids, int64_t* n_centroids,
                             fe_point_t* cluster_cloud, int64_t cap_cloud, int64_t* n_cloud);

/* estimateKeypoints, src:185-259: cropped cloud (intensity = elevation) -> keypoints and
 * keypoint_cloud. */
int fe_estimate_keypoints(fe_ctx_t* ctx, const fe_point_t* cloud, int64_t n,
                          fe_point_t* keypoints, int64_t cap_keypoints, int64_t* n_keypoints,
                          fe_point_t* keypoint_cloud, int64_t cap_cloud, int64_t* n_cloud);

/* estimateDescriptors, src:329-355: pcl::ShapeContext3DEstimation of `keypoints` against the
 * search surface `cloud_full`; descriptors is k x 1980 floats. */
int fe_estimate_descriptors(fe_ctx_t* ctx, const fe_point_t* cloud_full, int64_t n,
                            const fe_point_t* keypoints, int64_t k, float* descriptors);

/* concatenateFields, src:119 / feature_extraction_node.h:35-53: pack k keypoints and their
 * descriptors into pcl::PointDescriptor records of FE_RECORD_FLOATS floats (rf zeroed). */
int fe_pack_point_descriptors(const fe_point_t* keypoints, const float* descriptors, int64_t k,
                              float* records);

#ifdef __cplusplus
}
#endif
#endif /* FE_B200_H_ */

// core_example.cpp — the reference's cloudCallback body (src:83-117) written against the C++ host
// mirror, stage by stage and fused.  Build: see examples/Makefile.  Needs a B200 to run.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "feature_extraction_b200/feature_extraction_core.hpp"

using namespace feature_extraction_b200;

int main() {
  // a toy scan: one pole seen by rings 5..9, 7 points per ring, plus ground clutter
  PointCloud cloud_full;
  for (int ring = 5; ring <= 9; ring++) {
    const double el = ((ring - 7) * 2 - 1) * 3.14159265358979323846 / 180.0;
    for (int k = 0; k < 7; k++) {
      const double a = -1.0 + k / 3.0;
      Point p;
      p.x = (float)(10.0 - 0.08 * std::cos(a));
      p.y = (float)(2.0 + 0.08 * std::sin(a));
      p.z = (float)(std::hypot(p.x, p.y) * std::tan(el));
      p.intensity = 0.f;
      cloud_full.push_back(p);
    }
  }
  try {
    FeatureExtractionCore node(0);
    node.setImuRollPitch(3.14159265358979323846, 0.0);  // level: roll = tmproll - pi = 0

    // stage by stage, exactly the call sequence of cloudCallback
    PointCloud full = cloud_full;
    node.getElevationAngles(full);  // src:87
    node.rotateCloud(full);         // src:92
    PointCloud cloud = full;        // src:98
    node.filterCloud(cloud);        // src:99
    PointCloud keypoints, keypoint_cloud;
    node.estimateKeypoints(cloud, keypoints, keypoint_cloud);  // src:107
    DescriptorCloud descriptors;
    if (node.descriptorEstimation) node.estimateDescriptors(full, keypoints, descriptors);  // src:115

    // the same, fused
    PointCloud kp2;
    DescriptorCloud d2;
    node.processScan(cloud_full, kp2, d2);

    std::printf("stage-by-stage: %zu keypoints; fused: %zu keypoints\n", keypoints.size(), kp2.size());
    const bool same = keypoints.size() == kp2.size() &&
                      (keypoints.empty() || std::memcmp(keypoints.data(), kp2.data(), keypoints.size() * sizeof(Point)) == 0);
    for (size_t i = 0; i < kp2.size(); i++) std::printf("  kp %zu: %.4f %.4f %.4f el %.2f\n", i, kp2[i].x, kp2[i].y, kp2[i].z, kp2[i].intensity);
    // the ~features records (src:119) straight from the device against the host-side packing
    PointCloud kp3;
    std::vector<float> records, packed(kp2.size() * (size_t)FE_RECORD_FLOATS);
    node.processScanRecords(cloud_full, kp3, records);
    bool recSame = true;
    if (node.descriptorEstimation) {
      fe_pack_point_descriptors(kp2.data(), d2.data(), (int64_t)kp2.size(), packed.data());
      recSame = records.size() == packed.size() &&
                (packed.empty() || std::memcmp(records.data(), packed.data(), packed.size() * sizeof(float)) == 0);
    }
    std::printf("%s\n", (same && recSame) ? "identical" : "DIFFERENT");
    return (same && recSame) ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
}

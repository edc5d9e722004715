// check_glibc_f32 — compares feature_extraction_b200/csrc/glibc_f32.h (the restatement the device runs)
// with the libm this program links against, bit for bit.  Test infrastructure (tests/test_glibc_f32.py).
//   check_glibc_f32 <stride> <pairs>    acosf / atanf over every stride-th float, atan2f over <pairs> pairs
// Exit code 0 = identical everywhere.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../../feature_extraction_b200/csrc/glibc_f32.h"

using namespace fe::glibc;

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline bool same(float a, float b) { return bits(a) == bits(b) || (a != a && b != b); }

int main(int argc, char** argv) {
  const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  const long long pairs = argc > 2 ? atoll(argv[2]) : 1000000000LL;
  const int T = (int)std::max(1u, std::thread::hardware_concurrency());
  std::atomic<long long> bad_acos(0), bad_atan(0), bad_atan2(0), n_acos(0), n_atan(0), n_atan2(0);
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t]() {
      long long c = 0;
      for (uint64_t u = (uint64_t)t * stride; u <= 0xffffffffull; u += (uint64_t)T * stride, c++) {
        const float x = from_bits((uint32_t)u);
        volatile float vx = x;
        if (!same(acosf(vx), acosf_fdlibm(x)) && bad_acos++ < 5) printf("acosf(%a): libm %a, restatement %a\n", x, acosf(vx), acosf_fdlibm(x));
        if (!same(atanf(vx), atanf_fdlibm(x)) && bad_atan++ < 5) printf("atanf(%a): libm %a, restatement %a\n", x, atanf(vx), atanf_fdlibm(x));
      }
      n_acos += c; n_atan += c;
      // every float within 4096 ulps of the cosine of a 3DSC elevation edge (k*180/11 deg) and of +-1, 0, +-0.5
      for (int k = 0; k <= 11 && t == 0; k++) {
        const float cth = (float)cos((double)k * M_PI / 11.0);
        for (int d = -4096; d <= 4096; d++) {
          const float x = from_bits(bits(cth) + (uint32_t)d);
          volatile float vx = x;
          n_acos++;
          if (!same(acosf(vx), acosf_fdlibm(x)) && bad_acos++ < 5) printf("acosf(%a) near an edge: libm %a, restatement %a\n", x, acosf(vx), acosf_fdlibm(x));
        }
      }
      std::mt19937_64 g(1234 + t);
      std::uniform_real_distribution<float> U(-1.f, 1.f);
      const long long mine = pairs / T;
      for (long long i = 0; i < mine; i++) {
        float x, y;
        const int mode = (int)(i & 7);
        if (mode < 4) {  // what 3DSC passes: (|sin|, cos) of an azimuth, a near-unit vector
          const float ang = U(g) * 3.14159265f;
          x = cosf(ang); y = fabsf(sinf(ang));
          if (mode == 1) { const float s = 1.0f + U(g) * 1e-6f; x *= s; y *= s; }
        } else if (mode < 6) { x = U(g); y = fabsf(U(g)); }
        else if (mode == 6) { x = from_bits((uint32_t)g()); y = from_bits((uint32_t)g()); }
        else { x = U(g) * 1e-3f; y = fabsf(U(g)); }
        volatile float vx = x, vy = y;
        if (!same(atan2f(vy, vx), atan2f_fdlibm(y, x)) && bad_atan2++ < 5) printf("atan2f(%a, %a): libm %a, restatement %a\n", y, x, atan2f(vy, vx), atan2f_fdlibm(y, x));
      }
      n_atan2 += mine;
    });
  for (auto& t : th) t.join();
  // the azimuth edges themselves: phi = k*30 deg
  for (int k = 0; k <= 12; k++)
    for (int d = -4096; d <= 4096; d++) {
      const double a = (double)k * M_PI / 6.0;
      const float x = from_bits(bits((float)cos(a)) + (uint32_t)d), y = fabsf((float)sin(a));
      volatile float vx = x, vy = y;
      n_atan2++;
      if (!same(atan2f(vy, vx), atan2f_fdlibm(y, x)) && bad_atan2++ < 5) printf("atan2f(%a, %a) near an edge: libm %a, restatement %a\n", y, x, atan2f(vy, vx), atan2f_fdlibm(y, x));
    }
  printf("acosf: %lld values, %lld differ\natanf: %lld values, %lld differ\natan2f: %lld pairs, %lld differ\n", (long long)n_acos,
         (long long)bad_acos, (long long)n_atan, (long long)bad_atan, (long long)n_atan2, (long long)bad_atan2);
  return (bad_acos || bad_atan || bad_atan2) ? 1 : 0;
}

// rt_oracle.hpp — TEST INFRASTRUCTURE ONLY.  CPU f64 restatement of the reference's ray_color hot path
// (OrientalHorizon/Ray-Tracer-Archive, raytracer/src/*.rs).  Nothing in the product library may include,
// link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors, seeds or fixtures (cargo test runs zero tests,
// .github/workflows/run.yml:20-23; RNG is OS-seeded, rt_weekend.rs:8-15) and cannot be compiled here (no rustc/cargo,
// crates not vendored).  The oracle is pinned only by closed-form known answers (tests/test_oracle_kat.py) and by
// following the reference source line by line; every function cites the lines it restates.
#pragma once
#include <cstdint>
#include <cmath>

namespace orc {

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() = default;
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return (1.0 / s) * a; }  // vec3.rs Div: multiplies by 1/t
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vec3.rs:64-66
inline V3 cross(V3 u, V3 v) {                                                // vec3.rs:68-76
  return {u.y * v.z - u.z * v.y, -(u.x * v.z - u.z * v.x), u.x * v.y - u.y * v.x};
}
inline double length_squared(V3 a) { return dot(a, a); }
inline double length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 unit(V3 a) { return a / length(a); }                               // vec3.rs:29
inline V3 reflect(V3 v, V3 n) { return v - 2.0 * dot(v, n) * n; }            // vec3.rs:115-117
inline V3 refract(V3 uv, V3 n, double etai_over_etat) {                      // vec3.rs:246-251
  double cos_theta = std::fmin(dot(-uv, n), 1.0);
  V3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
  V3 r_out_parallel = -std::sqrt(std::fabs(1.0 - length_squared(r_out_perp))) * n;
  return r_out_perp + r_out_parallel;
}

struct Ray {  // ray.rs:1-39 — direction is NOT normalised
  V3 o, d;
  double tm = 0;
  V3 at(double t) const { return o + t * d; }
};

// ---- counter-based RNG shared with the device (replaces rand::random, rt_weekend.rs:8-19; only the
// distribution U[0,1) matters because the reference is OS-seeded).  Philox4x32-10, key = (pixel, sample),
// counter = (block, bounce, seed, 'RTB2').
struct Philox {
  static inline void round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  static inline void gen(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t bounce, uint32_t seed,
                         uint32_t out[4]) {
    uint32_t c[4] = {block, bounce, seed, 0x52544232u};
    uint32_t k[2] = {pixel, sample};
    for (int i = 0; i < 10; ++i) {
      round(c, k);
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
  static inline double u01(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }  // same 24 bits as the device
};

enum RngBlock : uint32_t { BLK_CAMERA0 = 0, BLK_CAMERA1 = 1, BLK_SCATTER = 2, BLK_AUX = 3, BLK_MEDIUM0 = 8 };

struct PathRng {
  uint32_t pixel = 0, sample = 0, seed = 0;
  void block(uint32_t blk, uint32_t bounce, double u[4]) const {
    uint32_t r[4];
    Philox::gen(pixel, sample, blk, bounce, seed, r);
    for (int i = 0; i < 4; ++i) u[i] = Philox::u01(r[i]);
  }
};

}  // namespace orc

// emul_traverse.cpp — TEST INFRASTRUCTURE.  Compiles the UNMODIFIED device header csrc/rtb_device.cuh with g++ by
// supplying host stand-ins for the handful of CUDA intrinsics it uses, so the exact traversal / intersection source
// the GPU runs can be checked against the oracle on a machine without a GPU.  Not linked into librtb200.so.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>  // float3/float4/uint4 + make_* ; __device__ etc. degrade to ignored attributes

static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
#undef __host__
#define __host__
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31u;
  return sh ? (hi << sh) | (lo >> (32u - sh)) : hi;
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  uint8_t b[8];
  for (int i = 0; i < 4; ++i) { b[i] = (x >> (8 * i)) & 0xFF; b[4 + i] = (y >> (8 * i)) & 0xFF; }
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)b[(s >> (4 * i)) & 7] << (8 * i);
  return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __any_sync(unsigned, int p) { return p; }  // (one 'lane': the warp-cooperative helpers are not run here)
static inline void __syncwarp() {}
#undef __forceinline__
#define __forceinline__ inline

#include "../../ray_tracer_archive_b200/csrc/rtb_device.cuh"

using namespace rtb;
static unsigned long long g_exact_rays = 0, g_refined_rays = 0;

#ifdef RTB_EMUL_STATS
static const float* g_tmax0 = nullptr;
#endif
extern "C" int emul_trace(const void* nodes, uint32_t n_nodes, const float* geom0, const uint32_t* info0,
                          const float* geom1, const uint32_t* info1, const float* geom2, const uint32_t* info2,
                          const float* geom3, const uint32_t* info3, const double* exact0, const double* exact1,
                          const double* exact2, float coord_max, float eps_ab, uint32_t global_f64, const uint32_t* globals, uint32_t n_globals,
                          uint32_t tree_empty, uint32_t n_snodes, const float* org,
                          const float* dir, const float* time, uint32_t n, uint32_t* ids, float* ts,
                          uint64_t* nodes_visited, uint64_t* prims_tested) {
  static DevScene sc;
  std::memset(&sc, 0, sizeof(sc));
  sc.nodes = (const uint4*)nodes;
  sc.n_nodes = n_nodes;
  sc.prmt_magic = 0x43000000u;
  sc.n_global = n_globals;
  sc.tree_empty = tree_empty;
  for (uint32_t k = 0; k < n_globals && k < RTB_MAX_GLOBALS; ++k) sc.global_ref[k] = globals[k];
  const float* g[4] = {geom0, geom1, geom2, geom3};
  const uint32_t* inf[4] = {info0, info1, info2, info3};
  for (int t = 0; t < 4; ++t) { sc.geom[t] = (const float4*)g[t]; sc.info[t] = (const uint2*)inf[t]; }
  static ExactTab xt;
  std::memset(&xt, 0, sizeof(xt));
  xt.tri = (const float4*)geom3;
  xt.exact[0] = exact0; xt.exact[1] = exact1; xt.exact[2] = exact2;
  for (int t = 0; t < 4; ++t) xt.info[t] = (const uint2*)inf[t];
  sc.xtab = &xt;
  sc.coord_max = coord_max; sc.eps_ab = eps_ab; sc.global_f64 = global_f64;
  uint64_t nv_total = 0, nt_total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    Closest best;
    uint32_t nv = 0;
    TestCount nt{};
    const float3 ro = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), rd = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    // the hot path (f32, ambiguity detection), then — as k_fixup does on the device — the exact pass if it asked for one
    float lo = 0.f;
#ifdef RTB_EMUL_STATS
    rtb::g_emul_tmax0 = g_tmax0 ? g_tmax0[i] : INFINITY;
#endif
    const uint32_t fix = traverse<true>(sc, (const uint4*)nodes, 0u, n_snodes < n_nodes ? n_snodes : n_nodes, ro, rd,
                                        time ? time[i] : 0.f, RTB_TMIN, best, lo, nv, nt);
    if (fix == FIX_RETRACE) {
      best = traverse_exact(sc, ro, rd, time ? time[i] : 0.f, lo, best.hi);
      ++g_exact_rays;
    } else if (fix == FIX_REFINE) {
      best.t = refine_hit(sc, best.ref, ro, rd, time ? time[i] : 0.f, best.t);
      ++g_refined_rays;
    }
    ids[i] = best.ref == REF_MISS ? RTB_NONE : ref_gid(sc, best.ref);
    ts[i] = best.t;
    nv_total += nv; nt_total += nt.n[0] + nt.n[1] + nt.n[2] + nt.n[3];
  }
  if (nodes_visited) *nodes_visited = nv_total;
  if (prims_tested) *prims_tested = nt_total;
  return 0;
}

// rays that went through the exact pass since the last call
extern "C" unsigned long long emul_exact_rays() {
  const unsigned long long n = g_exact_rays;
  g_exact_rays = 0;
  return n;
}
extern "C" unsigned long long emul_refined_rays() {
  const unsigned long long n = g_refined_rays;
  g_refined_rays = 0;
  return n;
}

// path numbering of the slot-stable pool (rtb_device.cuh: chunk_path), exposed for tests/test_pool_numbering.py
extern "C" unsigned long long emul_chunk_path(unsigned long long m, uint32_t chunk, uint32_t n_chunks) {
  return chunk_path(m, chunk, n_chunks);
}

#ifdef RTB_EMUL_STATS
extern "C" void emul_set_tmax0(const float* per_ray) { g_tmax0 = per_ray; }
extern "C" void emul_stats(unsigned long long* out8, int reset) {
  for (int i = 0; i < 8; ++i) { out8[i] = rtb::g_emul_stats[i]; if (reset) rtb::g_emul_stats[i] = 0; }
}
#endif

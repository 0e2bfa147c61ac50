// rt_oracle.cpp — TEST INFRASTRUCTURE ONLY (see rt_oracle.hpp header note; PARITY UNPINNED by the reference).
// Graph-mode oracle: interprets the same rtb_node records the product library consumes, but as the reference's
// trait-object world: nested HittableList linear scans, object-space Translate/RotateY wrappers, recursive-equivalent
// ray_color.  All arithmetic f64 (vec3.rs:7).
#include "rt_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <functional>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/rtb200.h"  // POD record layouts only (interface definition, no product code)

namespace orc {

static const double PI = 3.14159265358979323846;  // rt_weekend.rs:2
static const double INF = std::numeric_limits<double>::infinity();

struct HitRecord {  // hittable.rs:11-20
  V3 p, normal;
  int mat = -1;
  double t = 0, u = 0, v = 0;
  bool front_face = false;
  uint32_t prim = RTB_NONE;
  void set_face_normal(const Ray& r, V3 outward) {  // hittable.rs:41-48
    front_face = dot(r.d, outward) < 0.0;
    normal = front_face ? outward : -outward;
  }
};

struct HitCtx {  // carries the per-path stream for ConstantMedium's draw (constant_medium.rs:57)
  const PathRng* rng = nullptr;
  uint32_t bounce = 0;
};

struct Hittable {  // hittable.rs:51-60
  virtual ~Hittable() {}
  virtual bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) const = 0;
  virtual double pdf_value(V3, V3) const { return 0.0; }
  virtual V3 random(V3, double, double) const { return V3(1, 0, 0); }
};
typedef std::shared_ptr<Hittable> HP;

struct Sphere : Hittable {  // sphere.rs
  V3 c; double r; int mat; uint32_t id;
  static void uv(V3 p, double& u, double& v) {  // sphere.rs:32-37
    double theta = std::acos(-p.y);
    double phi = std::atan2(-p.z, p.x) + PI;
    u = phi / (2.0 * PI);
    v = theta / PI;
  }
  bool hit(const Ray& ray, double t_min, double t_max, HitRecord& rec, const HitCtx&) const override {  // sphere.rs:41-65
    V3 oc = ray.o - c;
    double a = length_squared(ray.d);
    double half_b = dot(oc, ray.d);
    double cc = length_squared(oc) - r * r;
    double det = half_b * half_b - a * cc;
    if (det < 0.0) return false;
    double sqrtd = std::sqrt(det);
    double root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
      root = (-half_b + sqrtd) / a;
      if (root < t_min || t_max < root) return false;
    }
    rec.t = root;
    rec.p = ray.at(root);
    V3 outward = (rec.p - c) / r;
    rec.set_face_normal(ray, outward);
    uv(outward, rec.u, rec.v);
    rec.mat = mat;
    rec.prim = id;
    return true;
  }
  double pdf_value(V3 o, V3 v) const override {  // sphere.rs:75-84
    HitRecord rec;
    Ray ray{o, v, 0.0};
    if (!hit(ray, 0.001, INF, rec, HitCtx())) return 0.0;
    double cos_theta_max = std::sqrt(1.0 - r * r / length_squared(c - o));
    double solid_angle = 2.0 * PI * (1.0 - cos_theta_max);
    return 1.0 / solid_angle;
  }
  V3 random(V3 o, double r1, double r2) const override;  // sphere.rs:85-90
};

struct Onb {  // onb.rs:19-42
  V3 u, v, w;
  explicit Onb(V3 n) {
    w = unit(n);
    V3 a = std::fabs(w.x) > 0.9 ? V3(0, 1, 0) : V3(1, 0, 0);
    v = unit(cross(w, a));
    u = cross(w, v);
  }
  V3 local(V3 a) const { return a.x * u + a.y * v + a.z * w; }
};

static V3 random_to_sphere(double radius, double distance_sq, double r1, double r2) {  // pdf.rs:82-91
  double z = 1.0 + r2 * (std::sqrt(1.0 - radius * radius / distance_sq) - 1.0);
  double phi = 2.0 * PI * r1;
  double x = std::cos(phi) * std::sqrt(1.0 - z * z);
  double y = std::sin(phi) * std::sqrt(1.0 - z * z);
  return V3(x, y, z);
}
V3 Sphere::random(V3 o, double r1, double r2) const {
  V3 direction = c - o;
  double distance_sq = length_squared(direction);
  Onb uvw(direction);
  return uvw.local(random_to_sphere(r, distance_sq, r1, r2));
}

struct MovingSphere : Hittable {  // moving_sphere.rs
  V3 c0, c1; double t0, t1, r; int mat; uint32_t id;
  V3 center(double time) const { return c0 + ((time - t0) / (t1 - t0)) * (c1 - c0); }  // :36-39
  bool hit(const Ray& ray, double t_min, double t_max, HitRecord& rec, const HitCtx&) const override {  // :43-66
    V3 oc = ray.o - center(ray.tm);
    double a = length_squared(ray.d);
    double half_b = dot(oc, ray.d);
    double cc = length_squared(oc) - r * r;
    double det = half_b * half_b - a * cc;
    if (det < 0.0) return false;
    double sqrtd = std::sqrt(det);
    double root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
      root = (-half_b + sqrtd) / a;
      if (root < t_min || t_max < root) return false;
    }
    rec.t = root;
    rec.p = ray.at(root);
    V3 outward = (rec.p - center(ray.tm)) / r;
    rec.set_face_normal(ray, outward);
    rec.u = 0; rec.v = 0;  // reference leaves u,v stale (:60-64); defined as 0
    rec.mat = mat;
    rec.prim = id;
    return true;
  }
};

// axis-aligned rects, aarect.rs.  axis = normal axis (2: XyRect, 1: XzRect, 0: YzRect); (ia, ib) = in-plane axes
struct AARect : Hittable {
  int axis, ia, ib; double a0, a1, b0, b1, k; int mat; uint32_t id;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx&) const override {  // aarect.rs:31-48,81-98,150-167
    double t = (k - r.o[axis]) / r.d[axis];
    if (t < t_min || t > t_max) return false;
    double a = r.o[ia] + t * r.d[ia];
    double b = r.o[ib] + t * r.d[ib];
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    rec.u = (a - a0) / (a1 - a0);
    rec.v = (b - b0) / (b1 - b0);
    rec.t = t;
    V3 n(axis == 0 ? 1 : 0, axis == 1 ? 1 : 0, axis == 2 ? 1 : 0);
    rec.set_face_normal(r, n);
    rec.mat = mat;
    rec.p = r.at(t);
    rec.prim = id;
    return true;
  }
  double pdf_value(V3 origin, V3 v) const override {  // only XzRect: aarect.rs:107-117
    if (axis != 1) return 0.0;
    HitRecord rec;
    Ray ray{origin, v, 0.0};
    if (!hit(ray, 0.001, INF, rec, HitCtx())) return 0.0;
    double area = (a1 - a0) * (b1 - b0);
    double distance_squared = rec.t * rec.t * length_squared(v);
    double cosine = std::fabs(dot(v, rec.normal) / length(v));
    return distance_squared / cosine / area;
  }
  V3 random(V3 origin, double r1, double r2) const override {  // only XzRect: aarect.rs:118-125
    if (axis != 1) return V3(1, 0, 0);
    V3 p(a0 + (a1 - a0) * r1, k, b0 + (b1 - b0) * r2);
    return p - origin;
  }
};

struct Triangle : Hittable {  // new: Moller-Trumbore with the reference's conventions
  V3 v0, e1, e2; int mat; uint32_t id;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx&) const override {
    V3 pv = cross(r.d, e2);
    double det = dot(e1, pv);
    if (det == 0.0) return false;
    double inv = 1.0 / det;
    V3 tv = r.o - v0;
    double u = dot(tv, pv) * inv;
    if (u < 0.0 || u > 1.0) return false;
    V3 qv = cross(tv, e1);
    double v = dot(r.d, qv) * inv;
    if (v < 0.0 || u + v > 1.0) return false;
    double t = dot(e2, qv) * inv;
    if (t < t_min || t > t_max) return false;  // closed range like aarect.rs:33
    rec.t = t; rec.u = u; rec.v = v;
    rec.p = r.at(t);
    rec.set_face_normal(r, unit(cross(e1, e2)));
    rec.mat = mat; rec.prim = id;
    return true;
  }
};

struct Quad : Hittable {  // new: RTTNW quad(Q,u,v); (alpha,beta) play the role of aarect's (u,v)
  V3 Q, u, v; int mat; uint32_t id;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx&) const override {
    V3 n = cross(u, v);
    double nn = dot(n, n);
    double denom = dot(n, r.d);
    if (denom == 0.0) return false;
    double t = dot(n, Q - r.o) / denom;
    if (t < t_min || t > t_max) return false;
    V3 pl = r.at(t) - Q;
    double alpha = dot(n, cross(pl, v)) / nn;
    double beta = dot(n, cross(u, pl)) / nn;
    if (alpha < 0.0 || alpha > 1.0 || beta < 0.0 || beta > 1.0) return false;
    rec.t = t; rec.u = alpha; rec.v = beta;
    rec.p = r.at(t);
    rec.set_face_normal(r, n / std::sqrt(nn));
    rec.mat = mat; rec.prim = id;
    return true;
  }
};

struct HittableList : Hittable {  // hittable_list.rs
  std::vector<HP> objects;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) const override {  // :39-51
    HitRecord temp;
    bool hit_anything = false;
    double closest = t_max;
    for (const HP& o : objects) {
      if (o->hit(r, t_min, closest, temp, cx)) {
        hit_anything = true;
        closest = temp.t;
        rec = temp;
      }
    }
    return hit_anything;
  }
  double pdf_value(V3 o, V3 v) const override {  // :73-80
    double weight = 1.0 / (double)objects.size();
    double sum = 0.0;
    for (const HP& ob : objects) sum += weight * ob->pdf_value(o, v);
    return sum;
  }
};

struct Translate : Hittable {  // hittable.rs:62-97
  HP ptr; V3 offset;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) const override {  // :76-85
    Ray moved{r.o - offset, r.d, r.tm};
    if (!ptr->hit(moved, t_min, t_max, rec, cx)) return false;
    rec.p = rec.p + offset;
    V3 norm = rec.normal;
    rec.set_face_normal(moved, norm);
    return true;
  }
};

struct RotateY : Hittable {  // hittable.rs:99-181
  HP ptr; double sin_theta, cos_theta;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) const override {  // :147-176
    V3 origin = r.o, direction = r.d;
    origin.x = cos_theta * r.o.x - sin_theta * r.o.z;
    origin.z = sin_theta * r.o.x + cos_theta * r.o.z;
    direction.x = cos_theta * r.d.x - sin_theta * r.d.z;
    direction.z = sin_theta * r.d.x + cos_theta * r.d.z;
    Ray rotated{origin, direction, r.tm};
    if (!ptr->hit(rotated, t_min, t_max, rec, cx)) return false;
    V3 p = rec.p, normal = rec.normal;
    p.x = cos_theta * rec.p.x + sin_theta * rec.p.z;
    p.z = -sin_theta * rec.p.x + cos_theta * rec.p.z;
    normal.x = cos_theta * rec.normal.x + sin_theta * rec.normal.z;
    normal.z = -sin_theta * rec.normal.x + cos_theta * rec.normal.z;
    rec.p = p;
    rec.set_face_normal(rotated, normal);  // (sic) object-space ray against world-space normal, :173
    return true;
  }
};

struct FlipFace : Hittable {  // hittable.rs:183-205
  HP ptr;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) const override {
    if (!ptr->hit(r, t_min, t_max, rec, cx)) return false;
    rec.front_face = !rec.front_face;
    return true;
  }
};

struct ConstantMedium : Hittable {  // constant_medium.rs:31-71 (commented in the reference; book-2 semantics)
  HP boundary; double neg_inv_density; int mat; uint32_t id; uint32_t medium_index;
  bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) const override {
    HitRecord rec1, rec2;
    if (!boundary->hit(r, -INF, INF, rec1, cx)) return false;
    if (!boundary->hit(r, rec1.t + 0.0001, INF, rec2, cx)) return false;
    if (rec1.t < t_min) rec1.t = t_min;
    if (rec2.t > t_max) rec2.t = t_max;
    if (rec1.t >= rec2.t) return false;
    if (rec1.t < 0.0) rec1.t = 0.0;
    double ray_length = length(r.d);
    double distance_inside = (rec2.t - rec1.t) * ray_length;
    double u[4] = {0.5, 0, 0, 0};
    if (cx.rng) cx.rng->block(BLK_MEDIUM0 + medium_index, cx.bounce, u);
    double hit_distance = neg_inv_density * std::log(u[0]);
    if (hit_distance > distance_inside) return false;
    rec.t = rec1.t + hit_distance / ray_length;
    rec.p = r.at(rec.t);
    rec.normal = V3(1, 0, 0);
    rec.front_face = true;
    rec.u = 0; rec.v = 0;
    rec.mat = mat;
    rec.prim = id;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------------
struct PerlinTable {  // perlin.rs:6-12; tables are generated by the caller (scene seed) and shared with the device
  V3 ranvec[256];
  uint32_t px[256], py[256], pz[256];
  double noise(V3 p) const {  // perlin.rs:26-52 — note the Hermite smoothing here AND again in perlin_interp
    double u = p.x - std::floor(p.x), v = p.y - std::floor(p.y), w = p.z - std::floor(p.z);
    u = u * u * (3.0 - 2.0 * u);
    v = v * v * (3.0 - 2.0 * v);
    w = w * w * (3.0 - 2.0 * w);
    int i = (int)std::floor(p.x), j = (int)std::floor(p.y), k = (int)std::floor(p.z);
    V3 c[2][2][2];
    for (int di = 0; di < 2; ++di)
      for (int dj = 0; dj < 2; ++dj)
        for (int dk = 0; dk < 2; ++dk)
          c[di][dj][dk] = ranvec[px[(i + di) & 255] ^ py[(j + dj) & 255] ^ pz[(k + dk) & 255]];
    // perlin_interp, perlin.rs:67-85
    double uu = u * u * (3.0 - 2.0 * u), vv = v * v * (3.0 - 2.0 * v), ww = w * w * (3.0 - 2.0 * w);
    double accum = 0.0;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int cc = 0; cc < 2; ++cc) {
          V3 weight_v(u - a, v - b, w - cc);
          accum += (a * uu + (1.0 - a) * (1.0 - uu)) * (b * vv + (1.0 - b) * (1.0 - vv)) *
                   (cc * ww + (1.0 - cc) * (1.0 - ww)) * dot(c[a][b][cc], weight_v);
        }
    return accum;
  }
  double turb(V3 p) const {  // perlin.rs:86-98
    double accum = 0.0, weight = 1.0;
    V3 tp = p;
    for (int i = 0; i < 7; ++i) {
      accum += weight * noise(tp);
      weight *= 0.5;
      tp = tp * 2.0;
    }
    return std::fabs(accum);
  }
};

struct Image { std::vector<uint8_t> data; uint32_t w = 0, h = 0; };

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }  // rt_weekend.rs:21-29

struct Scene {
  std::vector<rtb_node> nodes;
  std::vector<uint32_t> child_index;
  uint32_t root = 0;
  std::vector<rtb_material> mats;
  std::vector<rtb_texture> texs;
  std::vector<rtb_light> light_recs;
  std::vector<PerlinTable> perlins;
  std::vector<Image> images;
  struct Mesh { std::vector<float> verts; std::vector<uint32_t> idx; };
  std::vector<Mesh> meshes;
  // instantiated
  HP world;
  std::shared_ptr<HittableList> lights;
  uint32_t n_prims = 0, n_media = 0;
  bool built = false;
  std::string err;
  // ---- accelerated closest hit (optional): the SAME per-primitive reference tests, with the product's exported BVH used
  // only to cull candidates.  prim_obj[id] = primitive `id` re-wrapped in copies of its ancestor Translate / RotateY /
  // FlipFace wrappers, so prim_obj[id]->hit() is exactly what the nested list scan would evaluate for it.
  std::vector<HP> prim_obj;
  std::vector<HP> media_obj;
  std::vector<std::function<HP(HP)>> wrap_stack;
  struct Accel {
    std::vector<uint8_t> nodes;       // 80-byte nodes as exported by rtb_scene_export_bvh
    std::vector<uint32_t> info[4];    // (prim id, material|mode) pairs, leaf order, per type
    std::vector<uint32_t> globals;    // prim ids that are always candidates (kept out of the product's tree)
    bool on = false;
  } accel;

  HP wrapped(HP leaf) const {
    for (size_t k = wrap_stack.size(); k-- > 0;) leaf = wrap_stack[k](leaf);
    return leaf;
  }
  void register_prim(uint32_t id, HP leaf) {
    if (id == RTB_NONE) return;
    if (prim_obj.size() <= id) prim_obj.resize(id + 1);
    prim_obj[id] = wrapped(leaf);
  }

  V3 tex_value(uint32_t t, double u, double v, V3 p) const {
    const rtb_texture& tx = texs[t];
    switch (tx.type) {
      case RTB_TEX_SOLID: return V3(tx.rgb[0], tx.rgb[1], tx.rgb[2]);  // texture.rs:34-38
      case RTB_TEX_CHECKER: {                                          // texture.rs:60-69
        double sines = std::sin(10.0 * p.x) * std::sin(10.0 * p.y) * std::sin(10.0 * p.z);
        return sines < 0.0 ? tex_value(tx.odd, u, v, p) : tex_value(tx.even, u, v, p);
      }
      case RTB_TEX_NOISE: {                                            // texture.rs:90-96
        const PerlinTable& pt = perlins[tx.table];
        double s = 0.5 * (1.0 + std::sin(tx.scale * p.z + 10.0 * pt.turb(p)));
        return V3(s, s, s);
      }
      case RTB_TEX_IMAGE: {                                            // texture.rs:118-140
        if (tx.table == RTB_NONE || tx.table >= images.size() || images[tx.table].data.empty()) return V3(0, 1, 1);
        const Image& im = images[tx.table];
        u = clampd(u, 0.0, 1.0);
        v = 1.0 - clampd(v, 0.0, 1.0);
        uint32_t i = (uint32_t)(u * im.w), j = (uint32_t)(v * im.h);
        if (i >= im.w) i = im.w - 1;
        if (j >= im.h) j = im.h - 1;
        const double s = 1.0 / 255.0;
        size_t idx = ((size_t)j * im.w + i) * 3;
        return V3(s * im.data[idx], s * im.data[idx + 1], s * im.data[idx + 2]);
      }
    }
    return V3();
  }

  HP build_node(uint32_t ni, bool in_boundary) {
    if (ni >= nodes.size()) { err = "node index out of range"; return nullptr; }
    const rtb_node& n = nodes[ni];
    const double* p = n.p;
    auto child = [&](uint32_t k) -> HP {
      if (k >= n.n_children || n.first_child + k >= child_index.size()) { err = "missing child"; return nullptr; }
      return build_node(child_index[n.first_child + k], in_boundary);
    };
    auto new_id = [&]() -> uint32_t { return in_boundary ? RTB_NONE : n_prims++; };
    auto rect = [&](int axis, double a0, double a1, double b0, double b1, double k) -> HP {
      auto r = std::make_shared<AARect>();
      r->axis = axis;
      r->ia = axis == 0 ? 1 : 0;
      r->ib = axis == 2 ? 1 : 2;
      r->a0 = a0; r->a1 = a1; r->b0 = b0; r->b1 = b1; r->k = k;
      r->mat = (int)n.material;
      r->id = new_id();
      register_prim(r->id, r);
      return r;
    };
    switch (n.type) {
      case RTB_NODE_SPHERE: {
        auto s = std::make_shared<Sphere>();
        s->c = V3(p[0], p[1], p[2]); s->r = p[3]; s->mat = (int)n.material; s->id = new_id();
        register_prim(s->id, s);
        return s;
      }
      case RTB_NODE_MOVING_SPHERE: {
        auto s = std::make_shared<MovingSphere>();
        s->c0 = V3(p[0], p[1], p[2]); s->c1 = V3(p[3], p[4], p[5]);
        s->t0 = p[6]; s->t1 = p[7]; s->r = p[8]; s->mat = (int)n.material; s->id = new_id();
        register_prim(s->id, s);
        return s;
      }
      case RTB_NODE_XY_RECT: return rect(2, p[0], p[1], p[2], p[3], p[4]);
      case RTB_NODE_XZ_RECT: return rect(1, p[0], p[1], p[2], p[3], p[4]);
      case RTB_NODE_YZ_RECT: return rect(0, p[0], p[1], p[2], p[3], p[4]);
      case RTB_NODE_BOX: {  // boxes.rs:19-68, side order fixed
        auto l = std::make_shared<HittableList>();
        l->objects.push_back(rect(2, p[0], p[3], p[1], p[4], p[5]));
        l->objects.push_back(rect(2, p[0], p[3], p[1], p[4], p[2]));
        l->objects.push_back(rect(1, p[0], p[3], p[2], p[5], p[4]));
        l->objects.push_back(rect(1, p[0], p[3], p[2], p[5], p[1]));
        l->objects.push_back(rect(0, p[1], p[4], p[2], p[5], p[3]));
        l->objects.push_back(rect(0, p[1], p[4], p[2], p[5], p[0]));
        return l;
      }
      case RTB_NODE_TRIANGLE: {  // vertices are single precision by contract (include/rtb200.h), like mesh vertices
        auto t = std::make_shared<Triangle>();
        t->v0 = V3((float)p[0], (float)p[1], (float)p[2]);
        t->e1 = V3((float)p[3], (float)p[4], (float)p[5]) - t->v0;
        t->e2 = V3((float)p[6], (float)p[7], (float)p[8]) - t->v0;
        t->mat = (int)n.material; t->id = new_id();
        register_prim(t->id, t);
        return t;
      }
      case RTB_NODE_QUAD: {
        auto q = std::make_shared<Quad>();
        q->Q = V3(p[0], p[1], p[2]); q->u = V3(p[3], p[4], p[5]); q->v = V3(p[6], p[7], p[8]);
        q->mat = (int)n.material; q->id = new_id();
        register_prim(q->id, q);
        return q;
      }
      case RTB_NODE_MESH: {
        uint32_t mid = (uint32_t)p[0];
        if (mid >= meshes.size()) { err = "mesh id not set"; return nullptr; }
        const Mesh& m = meshes[mid];
        auto l = std::make_shared<HittableList>();
        for (size_t k = 0; k + 2 < m.idx.size(); k += 3) {
          auto t = std::make_shared<Triangle>();
          const float* a = &m.verts[3 * (size_t)m.idx[k]];
          const float* b = &m.verts[3 * (size_t)m.idx[k + 1]];
          const float* c = &m.verts[3 * (size_t)m.idx[k + 2]];
          t->v0 = V3(a[0], a[1], a[2]);
          t->e1 = V3(b[0], b[1], b[2]) - t->v0;
          t->e2 = V3(c[0], c[1], c[2]) - t->v0;
          t->mat = (int)n.material; t->id = new_id();
          register_prim(t->id, t);
          l->objects.push_back(t);
        }
        return l;
      }
      case RTB_NODE_TRANSLATE: {
        auto t = std::make_shared<Translate>();
        t->offset = V3(p[0], p[1], p[2]);
        const V3 off = t->offset;
        wrap_stack.push_back([off](HP c) -> HP { auto w = std::make_shared<Translate>(); w->ptr = c; w->offset = off; return w; });
        t->ptr = child(0);
        wrap_stack.pop_back();
        return t->ptr ? t : nullptr;
      }
      case RTB_NODE_ROTATE_Y: {  // hittable.rs:107-111
        auto r = std::make_shared<RotateY>();
        double radians = p[0] * PI / 180.0;
        r->sin_theta = std::sin(radians); r->cos_theta = std::cos(radians);
        const double sn = r->sin_theta, cs = r->cos_theta;
        wrap_stack.push_back([sn, cs](HP c) -> HP { auto w = std::make_shared<RotateY>(); w->ptr = c; w->sin_theta = sn; w->cos_theta = cs; return w; });
        r->ptr = child(0);
        wrap_stack.pop_back();
        return r->ptr ? r : nullptr;
      }
      case RTB_NODE_FLIP_FACE: {
        auto f = std::make_shared<FlipFace>();
        wrap_stack.push_back([](HP c) -> HP { auto w = std::make_shared<FlipFace>(); w->ptr = c; return w; });
        f->ptr = child(0);
        wrap_stack.pop_back();
        return f->ptr ? f : nullptr;
      }
      case RTB_NODE_CONSTANT_MEDIUM: {  // constant_medium.rs:23-29
        auto m = std::make_shared<ConstantMedium>();
        m->id = new_id();
        m->medium_index = n_media++;
        if (n.n_children < 1) { err = "medium without boundary"; return nullptr; }
        m->boundary = build_node(child_index[n.first_child], true);
        m->neg_inv_density = -1.0 / p[0];
        m->mat = (int)n.material;
        if (m->boundary && !in_boundary) media_obj.push_back(wrapped(m));
        return m->boundary ? m : nullptr;
      }
      case RTB_NODE_LIST:
      case RTB_NODE_BVH: {
        auto l = std::make_shared<HittableList>();
        for (uint32_t k = 0; k < n.n_children; ++k) {
          HP c = child(k);
          if (!c) return nullptr;
          l->objects.push_back(c);
        }
        return l;
      }
    }
    err = "unknown node type";
    return nullptr;
  }

  bool build() {
    if (built) return true;
    n_prims = 0; n_media = 0;
    prim_obj.clear(); media_obj.clear(); wrap_stack.clear();
    world = build_node(root, false);
    if (!world) return false;
    lights = std::make_shared<HittableList>();
    for (const rtb_light& l : light_recs) {  // the reference's separate, untransformed proxy list, main.rs:669-686
      if (l.type == RTB_LIGHT_XZ_RECT) {
        auto r = std::make_shared<AARect>();
        r->axis = 1; r->ia = 0; r->ib = 2;
        r->a0 = l.p[0]; r->a1 = l.p[1]; r->b0 = l.p[2]; r->b1 = l.p[3]; r->k = l.p[4];
        r->mat = -1; r->id = RTB_NONE;
        lights->objects.push_back(r);
      } else {
        auto s = std::make_shared<Sphere>();
        s->c = V3(l.p[0], l.p[1], l.p[2]); s->r = l.p[3]; s->mat = -1; s->id = RTB_NONE;
        lights->objects.push_back(s);
      }
    }
    built = true;
    return true;
  }
};


// Closest hit through the exported BVH: candidates from a plain f64 slab traversal of the (conservative, 7-bit
// quantised) child boxes, each candidate evaluated by the reference's own per-primitive test; the tie rule "equal t ->
// later primitive" (hittable_list.rs:44-47) is applied explicitly because candidates arrive in BVH order.
static bool hit_accel(const Scene& sc, const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx,
                      uint64_t* n_nodes, uint64_t* n_prims) {
  const uint8_t* base = sc.accel.nodes.data();
  const size_t n_nodes_total = sc.accel.nodes.size() / 80;
  if (n_nodes_total == 0) return false;
  double closest = t_max;
  bool found = false;
  uint32_t best_id = 0;
  HitRecord tmp;
  for (uint32_t gid : sc.accel.globals) {
    if (n_prims) ++*n_prims;
    if (sc.prim_obj[gid]->hit(r, t_min, closest, tmp, cx)) {
      if (!found || tmp.t < closest || gid > best_id) { found = true; closest = tmp.t; best_id = gid; rec = tmp; }
    }
  }
  uint32_t stack[256];
  int sp = 0;
  stack[sp++] = 0;
  const double inv[3] = {1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z};
  while (sp) {
    const uint32_t ni = stack[--sp];
    if (ni >= n_nodes_total) continue;
    const uint8_t* nd = base + (size_t)ni * 80;
    if (n_nodes) ++*n_nodes;
    float of[3];
    std::memcpy(of, nd, 12);
    const uint8_t* e = nd + 12;
    const uint32_t imask = nd[15];
    uint32_t child_base, prim_base;
    std::memcpy(&child_base, nd + 16, 4);
    std::memcpy(&prim_base, nd + 20, 4);
    const uint8_t* meta = nd + 24;
    const uint8_t* qlo = nd + 32;  // [3][8]
    const uint8_t* qhi = nd + 56;  // [3][8]
    uint32_t rank = 0;
    for (int s = 0; s < 8; ++s) {
      const bool internal = (imask >> s) & 1u;
      const uint32_t my_rank = rank;
      if (internal) ++rank;
      double tn = t_min, tf = closest;
      bool miss = false;
      for (int a = 0; a < 3 && !miss; ++a) {
        if (qlo[a * 8 + s] > qhi[a * 8 + s]) { miss = true; break; }  // empty slot
        const double step = std::ldexp(1.0, (int)e[a] - 127);
        const double ext = 127.0 * step;
        const double lo = (double)of[a] + qlo[a * 8 + s] * step - 1e-6 * ext;
        const double hi = (double)of[a] + qhi[a * 8 + s] * step + 1e-6 * ext;
        const double o = r.o[a], d = r.d[a];
        if (d == 0.0) {
          if (o < lo || o > hi) miss = true;
          continue;
        }
        double t0 = (lo - o) * inv[a], t1 = (hi - o) * inv[a];
        if (t0 > t1) std::swap(t0, t1);
        tn = std::fmax(tn, t0 - 1e-9 * std::fabs(t0));
        tf = std::fmin(tf, t1 + 1e-9 * std::fabs(t1));
        if (tn > tf) miss = true;
      }
      if (miss) continue;
      if (internal) {
        if (sp < 255) stack[sp++] = child_base + my_rank;
      } else {
        const uint32_t cnt = meta[s] >> 5, off = meta[s] & 31u;
        const uint32_t type = prim_base >> 29, first = (prim_base & ((1u << 29) - 1u)) + off;
        for (uint32_t k = 0; k < cnt; ++k) {
          const uint32_t gid = sc.accel.info[type][2 * (size_t)(first + k)];
          if (n_prims) ++*n_prims;
          if (sc.prim_obj[gid]->hit(r, t_min, closest, tmp, cx)) {
            if (!found || tmp.t < closest || gid > best_id) {
              found = true; closest = tmp.t; best_id = gid; rec = tmp;
            }
          }
        }
      }
    }
  }
  for (const HP& m : sc.media_obj) {  // ConstantMedium is order-independent: accepted iff its sampled t <= closest so far
    if (m->hit(r, t_min, closest, tmp, cx)) { found = true; closest = tmp.t; rec = tmp; }
  }
  return found;
}

static inline bool world_hit(const Scene& sc, const Ray& r, double t_min, double t_max, HitRecord& rec, const HitCtx& cx) {
  if (sc.accel.on) return hit_accel(sc, r, t_min, t_max, rec, cx, nullptr, nullptr);
  return sc.world->hit(r, t_min, t_max, rec, cx);
}

// ---- camera, camera.rs:21-70 -------------------------------------------------------------------------------
struct Camera {
  V3 origin, llc, horizontal, vertical, u, v, w;
  double lens_radius, time0, time1;
  explicit Camera(const rtb_camera& c) {
    double theta = c.vfov_deg * PI / 180.0;
    double h = std::tan(theta / 2.0);
    double vh = 2.0 * h, vw = c.aspect_ratio * vh;
    V3 from(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]), at(c.lookat[0], c.lookat[1], c.lookat[2]);
    V3 vup(c.vup[0], c.vup[1], c.vup[2]);
    w = unit(from - at);
    u = unit(cross(vup, w));
    v = cross(w, u);
    origin = from;
    horizontal = c.focus_dist * vw * u;
    vertical = c.focus_dist * vh * v;
    llc = origin - horizontal / 2.0 - vertical / 2.0 - c.focus_dist * w;
    lens_radius = c.aperture / 2.0;
    time0 = c.time0; time1 = c.time1;
  }
  // get_ray, camera.rs:60-70; the disk sample (dx,dy) and time are supplied by the caller's stream
  Ray get_ray(double s, double t, double dx, double dy, double time) const {
    V3 rd(lens_radius * dx, lens_radius * dy, 0);
    V3 offset = u * rd.x + v * rd.y;
    return Ray{origin + offset, llc + horizontal * s + vertical * t - origin - offset, time};
  }
};

// ---- the integrator: iterative form of ray_color, main.rs:63-139 ---------------------------------------------
struct SampleOut { V3 L; uint32_t segments; };

static SampleOut ray_color(const Scene& sc, Ray r, V3 background, int max_depth, uint32_t rr_start,
                           const PathRng& rng) {
  V3 L(0, 0, 0), beta(1, 1, 1);
  uint32_t segments = 0;
  const bool have_lights = !sc.lights->objects.empty();
  for (int depth = max_depth; depth > 0; --depth) {  // depth<=0 -> black, main.rs:71-73
    ++segments;
    const uint32_t bounce = segments;
    HitRecord rec;
    HitCtx cx; cx.rng = &rng; cx.bounce = bounce;
    if (!world_hit(sc, r, 0.001, INF, rec, cx)) {  // main.rs:74-76
      L = L + beta * background;
      break;
    }
    const rtb_material& m = sc.mats[rec.mat];
    double us[4], ua[4];
    rng.block(BLK_SCATTER, bounce, us);
    rng.block(BLK_AUX, bounce, ua);
    if (m.type == RTB_MAT_DIFFUSE_LIGHT) {  // emitted, material.rs:184-190; no scatter -> return emitted (main.rs:85-87)
      if (rec.front_face) L = L + beta * sc.tex_value(m.texture, rec.u, rec.v, rec.p);
      break;
    }
    if (m.type == RTB_MAT_METAL) {  // material.rs:95-107; specular branch main.rs:89-92
      double fuzz = m.param < 1.0 ? m.param : 1.0;
      V3 reflected = reflect(unit(r.d), rec.normal);
      // random_in_unit_sphere(): uniform BALL (vec3.rs:78-86), closed form: uniform direction * cbrt(xi)
      double z = 1.0 - 2.0 * ua[1], phi = 2.0 * PI * ua[2], rad = std::cbrt(ua[3]);
      double s = std::sqrt(std::fmax(0.0, 1.0 - z * z));
      V3 ball(rad * s * std::cos(phi), rad * s * std::sin(phi), rad * z);
      beta = beta * sc.tex_value(m.texture, rec.u, rec.v, rec.p);
      r = Ray{rec.p, reflected + fuzz * ball, 0.0};  // time reset to 0.0, material.rs:101
    } else if (m.type == RTB_MAT_DIELECTRIC) {  // material.rs:123-155
      double ratio = rec.front_face ? 1.0 / m.param : m.param;
      V3 ud = unit(r.d);
      double cos_theta = std::fmin(dot(-ud, rec.normal), 1.0);
      double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
      bool cannot_refract = ratio * sin_theta > 1.0;
      double r0 = (1.0 - ratio) / (1.0 + ratio);
      r0 *= r0;
      double reflectance = r0 + (1.0 - r0) * std::pow(1.0 - cos_theta, 5);  // material.rs:118-122
      V3 dir = (cannot_refract || reflectance > ua[1]) ? reflect(ud, rec.normal) : refract(ud, rec.normal, ratio);
      r = Ray{rec.p, dir, r.tm};
    } else {  // Lambertian (material.rs:48-71) or Isotropic (book-3 form): non-specular, main.rs:94-138
      const bool iso = m.type == RTB_MAT_ISOTROPIC;
      V3 atten = sc.tex_value(m.texture, rec.u, rec.v, rec.p);
      V3 dir;
      bool pick_light = have_lights && us[0] < 0.5;  // MixturePdf::generate, pdf.rs:73-79
      if (pick_light) {  // HittableList::random, hittable_list.rs:81-84 (uniform child)
        size_t n = sc.lights->objects.size();
        size_t k = (size_t)(us[1] * (double)n);
        if (k >= n) k = n - 1;
        dir = sc.lights->objects[k]->random(rec.p, us[2], us[3]);
      } else if (iso) {  // uniform sphere
        double z = 1.0 - 2.0 * us[2], phi = 2.0 * PI * us[3];
        double s = std::sqrt(std::fmax(0.0, 1.0 - z * z));
        dir = V3(s * std::cos(phi), s * std::sin(phi), z);
      } else {  // CosinePdf::generate, pdf.rs:32-34; random_cosine_direction vec3.rs:253-262 (r1 = us[2], r2 = us[3])
        double zz = std::sqrt(1.0 - us[3]);
        double phi = 2.0 * PI * us[2];
        V3 a(std::cos(phi) * std::sqrt(us[3]), std::sin(phi) * std::sqrt(us[3]), zz);
        dir = Onb(rec.normal).local(a);
      }
      double mat_pdf, spdf;
      if (iso) {
        mat_pdf = spdf = 1.0 / (4.0 * PI);
      } else {
        double cosine = dot(unit(dir), unit(rec.normal));  // CosinePdf::value pdf.rs:24-31 (w = unit(normal))
        mat_pdf = cosine <= 0.0 ? 0.0 : cosine / PI;
        double c2 = dot(rec.normal, unit(dir));            // scattering_pdf material.rs:64-71
        spdf = c2 < 0.0 ? 0.0 : c2 / PI;
      }
      double pdf_val = have_lights ? 0.5 * sc.lights->pdf_value(rec.p, dir) + 0.5 * mat_pdf : mat_pdf;  // pdf.rs:70-72
      if (!(spdf > 0.0)) break;  // zero-weight continuation culled, not a segment
      beta = beta * atten * (spdf / pdf_val);
      r = Ray{rec.p, dir, r.tm};
    }
    if (rr_start > 0 && segments >= rr_start) {  // Russian roulette (not in the reference; unbiased)
      double q = std::fmax(beta.x, std::fmax(beta.y, beta.z));
      q = q < 0.2 ? 0.2 : (q > 1.0 ? 1.0 : q);
      if (!(ua[0] < q)) break;
      beta = beta / q;
    }
  }
  return SampleOut{L, segments};
}

static void disk_sample(double u2, double u3, double& dx, double& dy) {  // uniform unit disk (vec3.rs:101-113), closed form
  double rr = std::sqrt(u2), phi = 2.0 * PI * u3;
  dx = rr * std::cos(phi);
  dy = rr * std::sin(phi);
}

}  // namespace orc

// ================================================= C API ====================================================
using namespace orc;

template <class F>
static void parallel_rows(uint32_t rows, int threads, F f) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  std::atomic<uint32_t> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&]() {
      for (;;) {
        uint32_t j = next.fetch_add(1);
        if (j >= rows) break;
        f(j);
      }
    });
  for (auto& th : pool) th.join();
}

extern "C" {

struct orc_scene { Scene s; };

orc_scene* orc_scene_create(const rtb_node* nodes, uint32_t n_nodes, const uint32_t* child_index, uint32_t n_child,
                            uint32_t root, const rtb_material* mats, uint32_t n_mats, const rtb_texture* texs,
                            uint32_t n_texs, const rtb_light* lights, uint32_t n_lights) {
  orc_scene* o = new orc_scene();
  o->s.nodes.assign(nodes, nodes + n_nodes);
  o->s.child_index.assign(child_index, child_index + n_child);
  o->s.root = root;
  o->s.mats.assign(mats, mats + n_mats);
  o->s.texs.assign(texs, texs + n_texs);
  if (n_lights) o->s.light_recs.assign(lights, lights + n_lights);
  return o;
}
void orc_scene_destroy(orc_scene* o) { delete o; }
const char* orc_scene_error(orc_scene* o) { return o->s.err.c_str(); }

int orc_scene_set_image(orc_scene* o, uint32_t id, const uint8_t* rgb, uint32_t w, uint32_t h) {
  if (o->s.images.size() <= id) o->s.images.resize(id + 1);
  o->s.images[id].data.assign(rgb, rgb + (size_t)w * h * 3);
  o->s.images[id].w = w; o->s.images[id].h = h;
  return 0;
}
int orc_scene_set_perlin(orc_scene* o, uint32_t id, const double* ranvec, const uint32_t* px, const uint32_t* py,
                         const uint32_t* pz) {
  if (o->s.perlins.size() <= id) o->s.perlins.resize(id + 1);
  PerlinTable& t = o->s.perlins[id];
  for (int i = 0; i < 256; ++i) {
    t.ranvec[i] = V3(ranvec[3 * i], ranvec[3 * i + 1], ranvec[3 * i + 2]);
    t.px[i] = px[i]; t.py[i] = py[i]; t.pz[i] = pz[i];
  }
  return 0;
}
int orc_scene_set_mesh(orc_scene* o, uint32_t id, const float* verts, uint32_t n_verts, const uint32_t* idx,
                       uint32_t n_tris) {
  if (o->s.meshes.size() <= id) o->s.meshes.resize(id + 1);
  o->s.meshes[id].verts.assign(verts, verts + (size_t)n_verts * 3);
  o->s.meshes[id].idx.assign(idx, idx + (size_t)n_tris * 3);
  o->s.built = false;
  return 0;
}
int orc_scene_build(orc_scene* o) { return o->s.build() ? 0 : -1; }
// attach the product's exported BVH (rtb_scene_export_bvh / rtb_scene_export_prims) as a candidate culler
int orc_scene_attach_bvh(orc_scene* o, const void* nodes80, uint32_t n_nodes, const uint32_t* info0, uint32_t n0,
                         const uint32_t* info1, uint32_t n1, const uint32_t* info2, uint32_t n2, const uint32_t* info3,
                         uint32_t n3, const uint32_t* global_ids, uint32_t n_globals) {
  if (!o->s.build()) return -1;
  Scene::Accel& a = o->s.accel;
  a.nodes.assign((const uint8_t*)nodes80, (const uint8_t*)nodes80 + (size_t)n_nodes * 80);
  const uint32_t* inf[4] = {info0, info1, info2, info3};
  const uint32_t cnt[4] = {n0, n1, n2, n3};
  for (int t = 0; t < 4; ++t) {
    a.info[t].clear();
    if (cnt[t]) a.info[t].assign(inf[t], inf[t] + (size_t)cnt[t] * 2);
    for (size_t k = 0; k < cnt[t]; ++k)
      if (a.info[t][2 * k] >= o->s.prim_obj.size() || !o->s.prim_obj[a.info[t][2 * k]]) { o->s.err = "BVH references an unknown primitive id"; return -2; }
  }
  a.globals.assign(global_ids, global_ids + n_globals);
  for (uint32_t g : a.globals)
    if (g >= o->s.prim_obj.size() || !o->s.prim_obj[g]) { o->s.err = "unknown global primitive id"; return -2; }
  a.on = true;
  return 0;
}
void orc_scene_use_bvh(orc_scene* o, int on) { o->s.accel.on = on && !o->s.accel.nodes.empty(); }
uint32_t orc_scene_num_prims(orc_scene* o) { return o->s.build() ? o->s.n_prims : 0; }

// pixel-centre primary rays (jitter 0.5, lens centre, time0); image row 0 = top = scanline j = H-1 (main.rs:733).
// stable_out / spread_out (optional): the device consumes f32 rays, so a pixel's answer is only well-defined if it
// survives an f32-ulp perturbation of the ray.  stable_out = 1 where the same primitive is returned for the ray with
// each direction component scaled by (1 +- eps); spread_out = max |t' - t| over those rays (the input-rounding
// uncertainty of t).  Rays through an exact edge (e.g. the Cornell box's x = y = 0 corner line seen from its
// symmetric camera) are knife-edge even in the reference's f64 and are reported as unstable rather than compared.
int orc_primary_hits(orc_scene* o, const rtb_camera* cam, uint32_t W, uint32_t H, uint32_t* ids, double* ts,
                     uint8_t* stable_out, double* spread_out, double eps, int threads) {
  if (!o->s.build()) return -1;
  Camera c(*cam);
  const Scene& sc = o->s;
  parallel_rows(H, threads, [&](uint32_t row) {
    uint32_t j = H - 1 - row;
    for (uint32_t i = 0; i < W; ++i) {
      double u = ((double)i + 0.5) / (double)(W - 1), v = ((double)j + 0.5) / (double)(H - 1);  // main.rs:752-753
      Ray r = c.get_ray(u, v, 0.0, 0.0, c.time0);
      HitRecord rec;
      bool h = world_hit(sc, r, 0.001, INF, rec, HitCtx());
      const size_t k = (size_t)row * W + i;
      ids[k] = h ? rec.prim : RTB_NONE;
      ts[k] = h ? rec.t : INF;
      if (stable_out) {
        bool stable = true;
        double spread = 0.0;
        for (int q = 0; q < 6 && stable; ++q) {
          Ray rp = r;
          rp.d.at(q >> 1) *= (q & 1) ? (1.0 + eps) : (1.0 - eps);
          HitRecord r2;
          bool h2 = world_hit(sc, rp, 0.001, INF, r2, HitCtx());
          if (h2 != h || (h && r2.prim != rec.prim)) stable = false;
          else if (h) spread = std::fmax(spread, std::fabs(r2.t - rec.t));
        }
        stable_out[k] = stable ? 1 : 0;
        if (spread_out) spread_out[k] = spread;
      }
    }
  });
  return 0;
}

int orc_trace_rays(orc_scene* o, const double* org, const double* dir, const double* time, uint32_t n, uint32_t* ids,
                   double* ts) {
  if (!o->s.build()) return -1;
  const uint32_t block = 4096;
  parallel_rows((n + block - 1) / block, n >= 8 * block ? 0 : 1, [&](uint32_t b) {
    const uint32_t end = std::min<uint64_t>(n, (uint64_t)(b + 1) * block);
    for (uint32_t i = b * block; i < end; ++i) {
      Ray r{V3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]),
            time ? time[i] : 0.0};
      HitRecord rec;
      bool h = world_hit(o->s, r, 0.001, INF, rec, HitCtx());
      ids[i] = h ? rec.prim : RTB_NONE;
      ts[i] = h ? rec.t : INF;
    }
  });
  return 0;
}

// accum: W*H*4 doubles (sum R, sum G, sum B, sum Y^2), row 0 = top.  Sample loop main.rs:751-762.
int orc_render(orc_scene* o, const rtb_camera* cam, const rtb_params* p, double* accum, uint64_t* segments_out,
               uint64_t* rejected_out, int threads) {
  if (!o->s.build()) return -1;
  Camera c(*cam);
  const Scene& sc = o->s;
  const uint32_t W = p->width, H = p->height;
  std::atomic<uint64_t> segs(0), rej(0);
  V3 bg(p->background[0], p->background[1], p->background[2]);
  parallel_rows(H, threads, [&](uint32_t row) {
    uint32_t j = H - 1 - row;
    uint64_t local_segs = 0, local_rej = 0;
    for (uint32_t i = 0; i < W; ++i) {
      double acc[4] = {0, 0, 0, 0};
      uint32_t pixel = row * W + i;
      for (uint32_t s = 0; s < p->spp; ++s) {
        PathRng rng; rng.pixel = pixel; rng.sample = p->sample_offset + s; rng.seed = p->seed;
        double u0[4], u1[4];
        rng.block(BLK_CAMERA0, 0, u0);
        rng.block(BLK_CAMERA1, 0, u1);
        double u = ((double)i + u0[0]) / (double)(W - 1), v = ((double)j + u0[1]) / (double)(H - 1);
        double dx, dy;
        disk_sample(u0[2], u0[3], dx, dy);
        double time = c.time0 + (c.time1 - c.time0) * u1[0];  // camera.rs:68
        Ray r = c.get_ray(u, v, dx, dy, time);
        SampleOut so = ray_color(sc, r, bg, p->max_depth, p->rr_start_depth, rng);
        local_segs += so.segments;
        if (!(std::isfinite(so.L.x) && std::isfinite(so.L.y) && std::isfinite(so.L.z))) { ++local_rej; continue; }
        double Y = 0.2126 * so.L.x + 0.7152 * so.L.y + 0.0722 * so.L.z;
        acc[0] += so.L.x; acc[1] += so.L.y; acc[2] += so.L.z; acc[3] += Y * Y;
      }
      double* dst = accum + (size_t)pixel * 4;
      for (int k = 0; k < 4; ++k) dst[k] = acc[k];
    }
    segs += local_segs; rej += local_rej;
  });
  if (segments_out) *segments_out = segs.load();
  if (rejected_out) *rejected_out = rej.load();
  return 0;
}

// write_color, main.rs:141-169 (operates on the per-pixel SUM)
void orc_write_color(const double* sum_rgb, uint32_t spp, uint8_t* out) {
  for (int k = 0; k < 3; ++k) {
    double c = sum_rgb[k];
    if (c != c) c = 0.0;
    c = std::sqrt((1.0 / (double)spp) * c);
    out[k] = (uint8_t)(256.0 * clampd(c, 0.0, 0.999));
  }
}

// ---- known-answer entry points (tests/test_oracle_kat.py) --------------------------------------------------
int orc_kat_sphere_hit(const double* c_r, const double* o, const double* d, double tmin, double tmax, double* out) {
  Sphere s; s.c = V3(c_r[0], c_r[1], c_r[2]); s.r = c_r[3]; s.mat = 0; s.id = 0;
  HitRecord rec;
  Ray r{V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), 0};
  if (!s.hit(r, tmin, tmax, rec, HitCtx())) return 0;
  out[0] = rec.t; out[1] = rec.normal.x; out[2] = rec.normal.y; out[3] = rec.normal.z;
  out[4] = rec.u; out[5] = rec.v; out[6] = rec.front_face ? 1 : 0;
  return 1;
}
int orc_kat_rect_hit(int axis, const double* abk, const double* o, const double* d, double tmin, double tmax,
                     double* out) {
  AARect a; a.axis = axis; a.ia = axis == 0 ? 1 : 0; a.ib = axis == 2 ? 1 : 2;
  a.a0 = abk[0]; a.a1 = abk[1]; a.b0 = abk[2]; a.b1 = abk[3]; a.k = abk[4]; a.mat = 0; a.id = 0;
  HitRecord rec;
  Ray r{V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), 0};
  if (!a.hit(r, tmin, tmax, rec, HitCtx())) return 0;
  out[0] = rec.t; out[1] = rec.normal.x; out[2] = rec.normal.y; out[3] = rec.normal.z;
  out[4] = rec.u; out[5] = rec.v; out[6] = rec.front_face ? 1 : 0;
  return 1;
}
void orc_kat_reflect(const double* v, const double* n, double* out) {
  V3 r = reflect(V3(v[0], v[1], v[2]), V3(n[0], n[1], n[2]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_kat_refract(const double* uv, const double* n, double eta, double* out) {
  V3 r = refract(V3(uv[0], uv[1], uv[2]), V3(n[0], n[1], n[2]), eta);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_kat_onb(const double* n, double* out9) {
  Onb b(V3(n[0], n[1], n[2]));
  out9[0] = b.u.x; out9[1] = b.u.y; out9[2] = b.u.z;
  out9[3] = b.v.x; out9[4] = b.v.y; out9[5] = b.v.z;
  out9[6] = b.w.x; out9[7] = b.w.y; out9[8] = b.w.z;
}
double orc_kat_sphere_pdf(const double* c_r, const double* o, const double* d) {
  Sphere s; s.c = V3(c_r[0], c_r[1], c_r[2]); s.r = c_r[3]; s.mat = 0; s.id = 0;
  return s.pdf_value(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]));
}
void orc_kat_sphere_random(const double* c_r, const double* o, double r1, double r2, double* out) {
  Sphere s; s.c = V3(c_r[0], c_r[1], c_r[2]); s.r = c_r[3]; s.mat = 0; s.id = 0;
  V3 d = s.random(V3(o[0], o[1], o[2]), r1, r2);
  out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
double orc_kat_xzrect_pdf(const double* abk, const double* o, const double* d) {
  AARect a; a.axis = 1; a.ia = 0; a.ib = 2;
  a.a0 = abk[0]; a.a1 = abk[1]; a.b0 = abk[2]; a.b1 = abk[3]; a.k = abk[4]; a.mat = 0; a.id = 0;
  return a.pdf_value(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]));
}
void orc_kat_philox(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t bounce, uint32_t seed, uint32_t* out4) {
  Philox::gen(pixel, sample, block, bounce, seed, out4);
}
double orc_kat_perlin_noise(orc_scene* o, uint32_t table, const double* p) {
  return o->s.perlins[table].noise(V3(p[0], p[1], p[2]));
}
double orc_kat_perlin_turb(orc_scene* o, uint32_t table, const double* p) {
  return o->s.perlins[table].turb(V3(p[0], p[1], p[2]));
}
void orc_kat_texture(orc_scene* o, uint32_t tex, double u, double v, const double* p, double* out) {
  V3 c = o->s.tex_value(tex, u, v, V3(p[0], p[1], p[2]));
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void orc_kat_camera_ray(const rtb_camera* cam, double s, double t, double dx, double dy, double time, double* out6) {
  Camera c(*cam);
  Ray r = c.get_ray(s, t, dx, dy, time);
  out6[0] = r.o.x; out6[1] = r.o.y; out6[2] = r.o.z; out6[3] = r.d.x; out6[4] = r.d.y; out6[5] = r.d.z;
}

}  // extern "C"

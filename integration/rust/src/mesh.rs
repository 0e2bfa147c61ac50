// raytracer/src/mesh.rs — NEW scene-construction types BASELINE.json's north_star names and the reference lacks
// (SURVEY §0, §8a N1): Triangle, general Quad, an indexed TriangleMesh and a Wavefront-OBJ loader (README.md:151-153 asks
// for one).  They live on the host side of the boundary: constructors + flatten() (flatten_impls.rs).  Their `hit` for
// the crate's own CPU path follows the reference's conventions (closed t-range like aarect.rs:33, set_face_normal with
// the geometric normal, uv = barycentrics / RTTNW alpha-beta).  NOT compiled in the build image.
use crate::aabb::Aabb;
use crate::hittable::{HitRecord, Hittable};
use crate::material::Material;
use crate::ray::Ray;
use crate::vec3::{cross, dot, Point3, Vec3};
use std::sync::Arc;

#[derive(Debug)]
pub struct Triangle {
    pub v0: Point3,
    pub v1: Point3,
    pub v2: Point3,
    pub mat_ptr: Arc<dyn Material>,
}
impl Triangle {
    pub fn construct(v0: &Point3, v1: &Point3, v2: &Point3, mat_ptr: Arc<dyn Material>) -> Self {
        // single precision by contract (the device tests triangles on f32 vertices)
        let r = |p: &Point3| Point3::construct(&[p.x() as f32 as f64, p.y() as f32 as f64, p.z() as f32 as f64]);
        Self { v0: r(v0), v1: r(v1), v2: r(v2), mat_ptr }
    }
}
impl Hittable for Triangle {
    fn hit(&self, r: &Ray, t_min: f64, t_max: f64, rec: &mut HitRecord) -> bool {
        // Moeller-Trumbore, the form the f64 oracle (oracle/rt_oracle.cpp: Triangle::hit) and the device's exact path use
        let (e1, e2) = (self.v1 - self.v0, self.v2 - self.v0);
        let pv = cross(&r.direction(), &e2);
        let det = dot(&e1, &pv);
        if det == 0.0 {
            return false;
        }
        let inv = 1.0 / det;
        let tv = r.origin() - self.v0;
        let u = dot(&tv, &pv) * inv;
        if u < 0.0 || u > 1.0 {
            return false;
        }
        let qv = cross(&tv, &e1);
        let v = dot(&r.direction(), &qv) * inv;
        if v < 0.0 || u + v > 1.0 {
            return false;
        }
        let t = dot(&e2, &qv) * inv;
        if t < t_min || t > t_max {
            return false;
        }
        rec.t = t;
        rec.u = u;
        rec.v = v;
        rec.p = r.at(t);
        rec.set_face_normal(r, &cross(&e1, &e2).unit());
        rec.mat_ptr = Some(Arc::clone(&self.mat_ptr));
        true
    }
    fn bounding_box(&self, _time0: f64, _time1: f64, output_box: &mut Aabb) -> bool {
        let lo = |a: f64, b: f64, c: f64| a.min(b).min(c) - 1e-4;
        let hi = |a: f64, b: f64, c: f64| a.max(b).max(c) + 1e-4;
        *output_box = Aabb::construct(
            &Point3::construct(&[lo(self.v0.x(), self.v1.x(), self.v2.x()), lo(self.v0.y(), self.v1.y(), self.v2.y()), lo(self.v0.z(), self.v1.z(), self.v2.z())]),
            &Point3::construct(&[hi(self.v0.x(), self.v1.x(), self.v2.x()), hi(self.v0.y(), self.v1.y(), self.v2.y()), hi(self.v0.z(), self.v1.z(), self.v2.z())]),
        );
        true
    }
}

#[derive(Debug)]
pub struct Quad {
    pub q: Point3,
    pub u: Vec3,
    pub v: Vec3,
    pub mat_ptr: Arc<dyn Material>,
}
impl Quad {
    pub fn construct(q: &Point3, u: &Vec3, v: &Vec3, mat_ptr: Arc<dyn Material>) -> Self {
        Self { q: *q, u: *u, v: *v, mat_ptr }
    }
}
impl Hittable for Quad {
    fn hit(&self, r: &Ray, t_min: f64, t_max: f64, rec: &mut HitRecord) -> bool {
        let n = cross(&self.u, &self.v);
        let nn = dot(&n, &n);
        let denom = dot(&n, &r.direction());
        if denom == 0.0 {
            return false;
        }
        let t = dot(&n, &(self.q - r.origin())) / denom;
        if t < t_min || t > t_max {
            return false;
        }
        let pl = r.at(t) - self.q;
        let alpha = dot(&n, &cross(&pl, &self.v)) / nn;
        let beta = dot(&n, &cross(&self.u, &pl)) / nn;
        if alpha < 0.0 || alpha > 1.0 || beta < 0.0 || beta > 1.0 {
            return false;
        }
        rec.t = t;
        rec.u = alpha;
        rec.v = beta;
        rec.p = r.at(t);
        rec.set_face_normal(r, &(n / nn.sqrt()));
        rec.mat_ptr = Some(Arc::clone(&self.mat_ptr));
        true
    }
    fn bounding_box(&self, _time0: f64, _time1: f64, output_box: &mut Aabb) -> bool {
        let c = [self.q, self.q + self.u, self.q + self.v, self.q + self.u + self.v];
        let mut lo = [f64::INFINITY; 3];
        let mut hi = [f64::NEG_INFINITY; 3];
        for p in &c {
            for a in 0..3 {
                lo[a] = lo[a].min(p.e[a] - 1e-4);
                hi[a] = hi[a].max(p.e[a] + 1e-4);
            }
        }
        *output_box = Aabb::construct(&Point3::construct(&lo), &Point3::construct(&hi));
        true
    }
}

/// Indexed triangle mesh: `vertices` = xyz triples (f32), `indices` = three per triangle.
#[derive(Debug)]
pub struct TriangleMesh {
    pub vertices: Vec<f32>,
    pub indices: Vec<u32>,
    pub mat_ptr: Arc<dyn Material>,
}
impl TriangleMesh {
    pub fn construct(vertices: Vec<f32>, indices: Vec<u32>, mat_ptr: Arc<dyn Material>) -> Self {
        assert!(indices.len() % 3 == 0 && indices.iter().all(|&i| (i as usize) * 3 + 2 < vertices.len()));
        Self { vertices, indices, mat_ptr }
    }
    /// Wavefront OBJ: `v x y z`, `f a b c ...` (1-based, negative = relative, `a/b/c` forms; polygons fan-triangulated)
    pub fn load_obj(text: &str, mat_ptr: Arc<dyn Material>, scale: f64, offset: &Vec3) -> Self {
        let (mut verts, mut idx) = (Vec::<f32>::new(), Vec::<u32>::new());
        for line in text.lines() {
            let mut it = line.split_whitespace();
            match it.next() {
                Some("v") => {
                    let c: Vec<f64> = it.take(3).map(|s| s.parse().expect("bad vertex")).collect();
                    for a in 0..3 {
                        verts.push((c[a] * scale + offset.e[a]) as f32);
                    }
                }
                Some("f") => {
                    let n = (verts.len() / 3) as i64;
                    let f: Vec<u32> = it
                        .map(|tok| {
                            let i: i64 = tok.split('/').next().unwrap().parse().expect("bad face index");
                            (if i > 0 { i - 1 } else { n + i }) as u32
                        })
                        .collect();
                    for k in 1..f.len().saturating_sub(1) {
                        idx.extend_from_slice(&[f[0], f[k], f[k + 1]]);
                    }
                }
                _ => {}
            }
        }
        Self::construct(verts, idx, mat_ptr)
    }
    fn tri(&self, k: usize) -> Triangle {
        let p = |i: u32| {
            let j = i as usize * 3;
            Point3::construct(&[self.vertices[j] as f64, self.vertices[j + 1] as f64, self.vertices[j + 2] as f64])
        };
        Triangle { v0: p(self.indices[3 * k]), v1: p(self.indices[3 * k + 1]), v2: p(self.indices[3 * k + 2]), mat_ptr: Arc::clone(&self.mat_ptr) }
    }
}
impl Hittable for TriangleMesh {
    // CPU path of the crate: linear scan in index order (= primitive-id order; later triangle wins equal t)
    fn hit(&self, r: &Ray, t_min: f64, t_max: f64, rec: &mut HitRecord) -> bool {
        let (mut any, mut closest) = (false, t_max);
        let mut tmp = HitRecord::new();
        for k in 0..self.indices.len() / 3 {
            if self.tri(k).hit(r, t_min, closest, &mut tmp) {
                any = true;
                closest = tmp.t;
                *rec = tmp.clone();
            }
        }
        any
    }
    fn bounding_box(&self, _time0: f64, _time1: f64, output_box: &mut Aabb) -> bool {
        let mut lo = [f64::INFINITY; 3];
        let mut hi = [f64::NEG_INFINITY; 3];
        for c in self.vertices.chunks(3) {
            for a in 0..3 {
                lo[a] = lo[a].min(c[a] as f64 - 1e-4);
                hi[a] = hi[a].max(c[a] as f64 + 1e-4);
            }
        }
        *output_box = Aabb::construct(&Point3::construct(&lo), &Point3::construct(&hi));
        !self.indices.is_empty()
    }
}

// raytracer/src/constant_medium.rs — the reference keeps this file entirely commented out, written against the book-2
// `scatter` signature (constant_medium.rs:1-75, material.rs:193-220).  BASELINE.json's configs 3 and 5 need it, so the shim
// restores it against the CURRENT traits (hittable.rs:51-60, material.rs:11-21) in the book-3 form SURVEY §8a M6 records:
// Isotropic is a non-specular material with the uniform-sphere pdf 1/(4 pi) (so it mixes with the light pdf in
// ray_color, main.rs:94-138).  `density` is kept next to neg_inv_density so that flatten() hands the library the
// constructor's own argument.  NOT compiled in the build image (no rustc / cargo).
use crate::aabb::Aabb;
use crate::hittable::{HitRecord, Hittable};
use crate::material::{Material, ScatterRecord};
use crate::pdf::Pdf;
use crate::ray::Ray;
use crate::rt_weekend::{random_double, INFINITY, PI};
use crate::texture::{SolidColor, Texture};
use crate::vec3::{random_unit_vector, Color3, Vec3};
use std::sync::Arc;

#[derive(Debug)]
pub struct Isotropic {
    pub albedo: Arc<dyn Texture>,
}
impl Isotropic {
    pub fn construct_color(albedo: &Color3) -> Self {
        Self { albedo: Arc::new(SolidColor::construct(albedo)) }
    }
}
/// uniform directions over the whole sphere
#[derive(Clone, Copy, Debug, Default)]
pub struct SpherePdf {}
impl Pdf for SpherePdf {
    fn value(&self, _direction: &Vec3) -> f64 {
        1.0 / (4.0 * PI)
    }
    fn generate(&self) -> Vec3 {
        random_unit_vector()
    }
}
impl Material for Isotropic {
    fn scatter(&self, _r_in: &Ray, rec: &HitRecord, srec: &mut ScatterRecord) -> bool {
        srec.is_specular = false;
        srec.attenuation = self.albedo.value(rec.u, rec.v, &rec.p);
        srec.pdf_ptr = Some(Arc::new(SpherePdf {}));
        true
    }
    fn scattering_pdf(&self, _r_in: &Ray, _rec: &HitRecord, _scattered: &Ray) -> f64 {
        1.0 / (4.0 * PI)
    }
}

#[derive(Debug)]
pub struct ConstantMedium {
    pub boundary: Arc<dyn Hittable>,
    pub phase_function: Arc<dyn Material>,
    pub density: f64,
    pub neg_inv_density: f64,
}
impl ConstantMedium {
    pub fn construct_color(b: Arc<dyn Hittable>, d: f64, c: &Color3) -> Self {
        Self { boundary: b, phase_function: Arc::new(Isotropic::construct_color(c)), density: d, neg_inv_density: -1.0 / d }
    }
}
impl Hittable for ConstantMedium {
    // the commented reference code (constant_medium.rs:31-71), statement for statement, on the current HitRecord
    fn hit(&self, r: &Ray, t_min: f64, t_max: f64, rec: &mut HitRecord) -> bool {
        let mut rec1 = HitRecord::new();
        let mut rec2 = HitRecord::new();
        if !self.boundary.hit(r, -INFINITY, INFINITY, &mut rec1) {
            return false;
        }
        if !self.boundary.hit(r, rec1.t + 0.0001, INFINITY, &mut rec2) {
            return false;
        }
        if rec1.t < t_min {
            rec1.t = t_min;
        }
        if rec2.t > t_max {
            rec2.t = t_max;
        }
        if rec1.t >= rec2.t {
            return false;
        }
        if rec1.t < 0.0 {
            rec1.t = 0.0;
        }
        let ray_length = r.direction().length();
        let distance_inside_boundary = (rec2.t - rec1.t) * ray_length;
        let hit_distance = self.neg_inv_density * random_double().ln();
        if hit_distance > distance_inside_boundary {
            return false;
        }
        rec.t = rec1.t + hit_distance / ray_length;
        rec.p = r.at(rec.t);
        rec.normal = Vec3::construct(&[1.0, 0.0, 0.0]); // arbitrary
        rec.front_face = true; // also arbitrary
        rec.mat_ptr = Some(Arc::clone(&self.phase_function));
        true
    }
    fn bounding_box(&self, time0: f64, time1: f64, output_box: &mut Aabb) -> bool {
        self.boundary.bounding_box(time0, time1, output_box)
    }
}

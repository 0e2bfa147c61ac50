// raytracer/src/flatten_impls.rs — `flatten()` for EVERY Hittable / Material / Texture impl of the reference: one record
// per constructor call, children before parents (post-order), exactly the records include/rtb200_scene.hpp emits.
//
// How it attaches (three one-line trait edits, three one-field struct additions; nothing else in the crate changes):
//   hittable.rs:51   pub trait Hittable: Debug + Send + Sync + FlattenHittable { ... }
//   material.rs:11   pub trait Material: Debug + Send + Sync + FlattenMaterial { ... }
//   texture.rs:7     pub trait Texture:  Debug + Send + Sync + FlattenTexture  { ... }
//   hittable.rs:99   RotateY gains `pub angle: f64` (degrees, set in construct(); the record carries the constructor's
//                    argument so that the library recomputes the very sin / cos of hittable.rs:108-110)
//   boxes.rs:11      Box gains `pub mat: Arc<dyn Material>` (set in construct(); `sides` only hold it behind dyn Hittable)
//   bvh.rs:10        BVHNode gains `pub src: Vec<Arc<dyn Hittable>>` (the list construct2() was given: the reference's
//                    tree drops ~30 % of the objects, bvh.rs:86,108 — SURVEY §0 — so the LIST is what gets flattened)
// NOT compiled in the build image (no rustc / cargo); tests/cpp/test_scene_mirror.cpp pushes the same records for
// cornell_box() and final_scene() through the same C ABI and renders them.
use crate::aarect::{XyRect, XzRect, YzRect};
use crate::boxes::Box;
use crate::bvh::BVHNode;
use crate::constant_medium::{ConstantMedium, Isotropic};
use crate::gpu::*;
use crate::hittable::{FlipFace, RotateY, Translate};
use crate::hittable_list::HittableList;
use crate::material::{Dielectric, DiffuseLight, Lambertian, Metal};
use crate::mesh::{Quad, Triangle, TriangleMesh};
use crate::moving_sphere::MovingSphere;
use crate::sphere::Sphere;
use crate::texture::{CheckerTexture, ImageTexture, NoiseTexture, SolidColor};
use std::sync::Arc;

pub trait FlattenHittable {
    /// pushes this object's record (after its children's) and returns its index
    fn flatten(&self, b: &mut SceneBuilder) -> u32;
}
pub trait FlattenMaterial {
    /// `key` = address of the Arc this material lives in: a shared material is recorded once
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32;
}
pub trait FlattenTexture {
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32;
}

fn mat_id(m: &Arc<dyn crate::material::Material>, b: &mut SceneBuilder) -> u32 {
    let key = Arc::as_ptr(m) as *const () as usize;
    if let Some(&id) = b.material_ids.get(&key) {
        return id;
    }
    m.flatten(key, b)
}
fn tex_id(t: &Arc<dyn crate::texture::Texture>, b: &mut SceneBuilder) -> u32 {
    let key = Arc::as_ptr(t) as *const () as usize;
    if let Some(&id) = b.texture_ids.get(&key) {
        return id;
    }
    t.flatten(key, b)
}
fn blank_texture(ty: u32) -> RtbTexture {
    RtbTexture { ty, even: RTB_NONE, odd: RTB_NONE, table: RTB_NONE, rgb: [0.0; 3], scale: 0.0 }
}

// ------------------------------------------------------------------------------------------------ textures
impl FlattenTexture for SolidColor {
    // texture.rs:12-38
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let mut t = blank_texture(TEX_SOLID);
        t.rgb = [self.color_value.x(), self.color_value.y(), self.color_value.z()];
        b.push_texture(key, t)
    }
}
impl FlattenTexture for CheckerTexture {
    // texture.rs:40-69: children first
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let mut t = blank_texture(TEX_CHECKER);
        t.even = tex_id(&self.even, b);
        t.odd = tex_id(&self.odd, b);
        b.push_texture(key, t)
    }
}
impl FlattenTexture for NoiseTexture {
    // texture.rs:71-96 + perlin.rs:6-25: the tables travel with the texture (the device evaluates perlin.rs:26-98 on them)
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let mut t = blank_texture(TEX_NOISE);
        t.scale = self.scale;
        t.table = b.perlins.len() as u32;
        let mut rv = Vec::with_capacity(768);
        for v in &self.noise.ranvec {
            rv.extend_from_slice(&[v.x(), v.y(), v.z()]);
        }
        b.perlins.push((rv, self.noise.perm_x.clone(), self.noise.perm_y.clone(), self.noise.perm_z.clone()));
        b.push_texture(key, t)
    }
}
impl FlattenTexture for ImageTexture {
    // texture.rs:98-141: raw RGB8 rows from the top; an empty image keeps table = RTB_NONE (renders cyan, :119-121)
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let mut t = blank_texture(TEX_IMAGE);
        if !self.data.is_empty() {
            t.table = b.images.len() as u32;
            b.images.push((self.data.clone(), self.width, self.height));
        }
        b.push_texture(key, t)
    }
}

// ------------------------------------------------------------------------------------------------ materials
impl FlattenMaterial for Lambertian {
    // material.rs:24-72
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let t = tex_id(&self.albedo, b);
        b.push_material(key, MAT_LAMBERTIAN, t, 0.0)
    }
}
impl FlattenMaterial for Metal {
    // material.rs:74-108: albedo is a plain colour -> its own solid texture; fuzz already clamped to <= 1 (:90-93)
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let mut t = blank_texture(TEX_SOLID);
        t.rgb = [self.albedo.x(), self.albedo.y(), self.albedo.z()];
        b.textures.push(t);
        let tid = (b.textures.len() - 1) as u32;
        b.push_material(key, MAT_METAL, tid, self.fuzz)
    }
}
impl FlattenMaterial for Dielectric {
    // material.rs:110-156
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        b.push_material(key, MAT_DIELECTRIC, RTB_NONE, self.ir)
    }
}
impl FlattenMaterial for DiffuseLight {
    // material.rs:158-191
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let t = tex_id(&self.emit, b);
        b.push_material(key, MAT_DIFFUSE_LIGHT, t, 0.0)
    }
}
impl FlattenMaterial for Isotropic {
    // material.rs:193-220 (commented out in the reference; restored in the shim's constant_medium.rs)
    fn flatten(&self, key: usize, b: &mut SceneBuilder) -> u32 {
        let t = tex_id(&self.albedo, b);
        b.push_material(key, MAT_ISOTROPIC, t, 0.0)
    }
}

// ------------------------------------------------------------------------------------------------ hittables
impl FlattenHittable for Sphere {
    // sphere.rs:12-38
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mat_ptr, b);
        b.leaf(NODE_SPHERE, m, &[self.center.x(), self.center.y(), self.center.z(), self.radius])
    }
}
impl FlattenHittable for MovingSphere {
    // moving_sphere.rs:9-34
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mat_ptr, b);
        b.leaf(
            NODE_MOVING_SPHERE,
            m,
            &[
                self.center0.x(), self.center0.y(), self.center0.z(),
                self.center1.x(), self.center1.y(), self.center1.z(),
                self.time0, self.time1, self.radius,
            ],
        )
    }
}
impl FlattenHittable for XyRect {
    // aarect.rs:10-29
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mp, b);
        b.leaf(NODE_XY_RECT, m, &[self.x0, self.x1, self.y0, self.y1, self.k])
    }
}
impl FlattenHittable for XzRect {
    // aarect.rs:60-79
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mp, b);
        b.leaf(NODE_XZ_RECT, m, &[self.x0, self.x1, self.z0, self.z1, self.k])
    }
}
impl FlattenHittable for YzRect {
    // aarect.rs:129-148
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mp, b);
        b.leaf(NODE_YZ_RECT, m, &[self.y0, self.y1, self.z0, self.z1, self.k])
    }
}
impl FlattenHittable for Box {
    // boxes.rs:11-75: ONE record; the library expands the six sides in the order of boxes.rs:19-68
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mat, b);
        b.leaf(
            NODE_BOX,
            m,
            &[self.box_min.x(), self.box_min.y(), self.box_min.z(), self.box_max.x(), self.box_max.y(), self.box_max.z()],
        )
    }
}
impl FlattenHittable for Translate {
    // hittable.rs:62-97
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let c = self.ptr.flatten(b);
        b.inner(NODE_TRANSLATE, RTB_NONE, &[self.offset.x(), self.offset.y(), self.offset.z()], &[c])
    }
}
impl FlattenHittable for RotateY {
    // hittable.rs:99-181: p[0] = the constructor's angle in degrees
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let c = self.ptr.flatten(b);
        b.inner(NODE_ROTATE_Y, RTB_NONE, &[self.angle], &[c])
    }
}
impl FlattenHittable for FlipFace {
    // hittable.rs:183-205
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let c = self.ptr.flatten(b);
        b.inner(NODE_FLIP_FACE, RTB_NONE, &[], &[c])
    }
}
impl FlattenHittable for ConstantMedium {
    // constant_medium.rs:8-29 (commented out in the reference): p[0] = density, material = the Isotropic phase function,
    // child = the boundary (a Sphere or a Box, optionally under Translate / RotateY; convex, as the reference requires)
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let c = self.boundary.flatten(b);
        let m = mat_id(&self.phase_function, b);
        b.inner(NODE_CONSTANT_MEDIUM, m, &[self.density], &[c])
    }
}
impl FlattenHittable for HittableList {
    // hittable_list.rs:12-37: children in list order = primitive-id order (the "later object wins equal t" rule, :44-47)
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let kids: Vec<u32> = self.objects.iter().map(|o| o.flatten(b)).collect();
        b.inner(NODE_LIST, RTB_NONE, &[], &kids)
    }
}
impl FlattenHittable for BVHNode {
    // bvh.rs:10-148: same closest-hit semantics as the list it was built from; the library builds its own wide BVH
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let kids: Vec<u32> = self.src.iter().map(|o| o.flatten(b)).collect();
        b.inner(NODE_BVH, RTB_NONE, &[], &kids)
    }
}
impl FlattenHittable for Triangle {
    // mesh.rs (new): vertices are single precision by contract
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mat_ptr, b);
        let (a, c, d) = (&self.v0, &self.v1, &self.v2);
        b.leaf(NODE_TRIANGLE, m, &[a.x(), a.y(), a.z(), c.x(), c.y(), c.z(), d.x(), d.y(), d.z()])
    }
}
impl FlattenHittable for Quad {
    // mesh.rs (new): RTTNW quad(Q, u, v)
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mat_ptr, b);
        let (q, u, v) = (&self.q, &self.u, &self.v);
        b.leaf(NODE_QUAD, m, &[q.x(), q.y(), q.z(), u.x(), u.y(), u.z(), v.x(), v.y(), v.z()])
    }
}
impl FlattenHittable for TriangleMesh {
    // mesh.rs (new): indexed mesh (OBJ loader output); triangles take consecutive primitive ids in index order
    fn flatten(&self, b: &mut SceneBuilder) -> u32 {
        let m = mat_id(&self.mat_ptr, b);
        let id = b.meshes.len() as f64;
        b.meshes.push((self.vertices.clone(), self.indices.clone()));
        b.leaf(NODE_MESH, m, &[id])
    }
}

/// the reference's separate light-proxy list (main.rs:669-686): only XzRect and Sphere implement pdf_value / random
/// (aarect.rs:107-125, sphere.rs:75-90), so only they can be handed to rtb_scene_set_lights
pub fn light_records(rects: &[&XzRect], spheres: &[&Sphere]) -> Vec<RtbLight> {
    let mut out = Vec::new();
    for r in rects {
        out.push(RtbLight { ty: LIGHT_XZ_RECT, _pad: 0, p: [r.x0, r.x1, r.z0, r.z1, r.k] });
    }
    for s in spheres {
        out.push(RtbLight { ty: LIGHT_SPHERE, _pad: 0, p: [s.center.x(), s.center.y(), s.center.z(), s.radius, 0.0] });
    }
    out
}

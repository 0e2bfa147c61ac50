// One new trait method per Hittable/Material/Texture impl (hittable.rs:51-60, material.rs:11-21, texture.rs:7-9).
// Sketch of the mechanical additions; see INTEGRATION.md.
pub trait Hittable: Debug + Send + Sync { /* hit, bounding_box, pdf_value, random as today */
    fn flatten(&self, b: &mut SceneBuilder) -> u32; }
impl Hittable for Sphere   { fn flatten(&self, b: &mut SceneBuilder) -> u32 { let m = self.mat_ptr.flatten(b);
    b.leaf(1 /*RTB_NODE_SPHERE*/, m, &[self.center.x(), self.center.y(), self.center.z(), self.radius]) } }
impl Hittable for XzRect   { fn flatten(&self, b: &mut SceneBuilder) -> u32 { let m = self.mp.flatten(b);
    b.leaf(4 /*RTB_NODE_XZ_RECT*/, m, &[self.x0, self.x1, self.z0, self.z1, self.k]) } }
impl Hittable for Box      { /* RTB_NODE_BOX, p = box_min ++ box_max: the library expands the 6 sides in boxes.rs:19-68 order */ }
impl Hittable for Translate{ fn flatten(&self, b: &mut SceneBuilder) -> u32 { let c = self.ptr.flatten(b);
    b.inner(16, u32::MAX, &[self.offset.x(), self.offset.y(), self.offset.z()], &[c]) } }
impl Hittable for RotateY  { /* 17, p = [angle in degrees] (store the angle in the struct next to sin/cos) */ }
impl Hittable for FlipFace { /* 18 */ }
impl Hittable for HittableList { fn flatten(&self, b: &mut SceneBuilder) -> u32 {
    let kids: Vec<u32> = self.objects.iter().map(|o| o.flatten(b)).collect(); b.inner(32, u32::MAX, &[], &kids) } }
// Material::flatten pushes an RtbMaterial (type 0..4, texture id, fuzz / ir); Texture::flatten an RtbTexture.

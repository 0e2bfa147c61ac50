// raytracer/src/gpu.rs — extern "C" declarations mirroring include/rtb200.h one to one, plus the record builder the
// `flatten()` methods push into.  NOT compiled in the build image (no rustc/cargo there); the ctypes binding
// (ray_tracer_archive_b200/_ffi.py) binds the same symbols with the same layouts and is what the tests exercise.
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy)]
pub struct RtbNode { pub ty: u32, pub material: u32, pub first_child: u32, pub n_children: u32, pub p: [f64; 12] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtbMaterial { pub ty: u32, pub texture: u32, pub param: f64 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtbTexture { pub ty: u32, pub even: u32, pub odd: u32, pub table: u32, pub rgb: [f64; 3], pub scale: f64 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtbLight { pub ty: u32, pub _pad: u32, pub p: [f64; 5] }
#[repr(C)] pub struct RtbCamera { pub lookfrom: [f64; 3], pub lookat: [f64; 3], pub vup: [f64; 3], pub vfov_deg: f64,
    pub aspect_ratio: f64, pub aperture: f64, pub focus_dist: f64, pub time0: f64, pub time1: f64 }
#[repr(C)] pub struct RtbParams { pub width: u32, pub height: u32, pub spp: u32, pub sample_offset: u32,
    pub total_spp: u32, pub max_depth: i32, pub rr_start_depth: u32, pub seed: u32, pub background: [f32; 3],
    pub pool_paths: u32, pub flags: u32 }
#[repr(C)] #[derive(Default)] pub struct RtbStats { pub paths: u64, pub segments: u64, pub rejected: u64,
    pub iterations: u64, pub launches: u64, pub extend_launches: u64, pub ms_total: f64, pub ms_extend: f64,
    pub nodes_visited: u64, pub prims_tested: u64 }
pub enum RtbContext {} pub enum RtbScene {}

extern "C" {
    pub fn rtb_last_error() -> *const c_char;
    pub fn rtb_context_create(device_id: c_int, out: *mut *mut RtbContext) -> c_int;
    pub fn rtb_context_destroy(ctx: *mut RtbContext);
    pub fn rtb_scene_create(ctx: *mut RtbContext, out: *mut *mut RtbScene) -> c_int;
    pub fn rtb_scene_destroy(s: *mut RtbScene);
    pub fn rtb_scene_set_materials(s: *mut RtbScene, m: *const RtbMaterial, n: u32) -> c_int;
    pub fn rtb_scene_set_textures(s: *mut RtbScene, t: *const RtbTexture, n: u32) -> c_int;
    pub fn rtb_scene_set_image(s: *mut RtbScene, id: u32, rgb: *const u8, w: u32, h: u32) -> c_int;
    pub fn rtb_scene_set_perlin(s: *mut RtbScene, id: u32, ranvec: *const f64, px: *const u32, py: *const u32, pz: *const u32) -> c_int;
    pub fn rtb_scene_set_lights(s: *mut RtbScene, l: *const RtbLight, n: u32) -> c_int;
    pub fn rtb_scene_set_graph(s: *mut RtbScene, nodes: *const RtbNode, n: u32, child: *const u32, nc: u32, root: u32) -> c_int;
    pub fn rtb_scene_commit(s: *mut RtbScene) -> c_int;
    pub fn rtb_render(ctx: *mut RtbContext, s: *mut RtbScene, cam: *const RtbCamera, p: *const RtbParams,
                      accum_out: *mut f32, stats: *mut RtbStats) -> c_int;
    pub fn rtb_finalize_rgb8(ctx: *mut RtbContext, d_accum: *const c_void, w: u32, h: u32, total_spp: u32, out: *mut u8) -> c_int;
}

/// Collects the records; ids of materials/textures are positions in these vectors.
#[derive(Default)]
pub struct SceneBuilder { pub nodes: Vec<RtbNode>, pub children: Vec<u32>, pub materials: Vec<RtbMaterial>,
    pub textures: Vec<RtbTexture>, pub images: Vec<(Vec<u8>, u32, u32)>, pub perlins: Vec<(Vec<f64>, Vec<u32>, Vec<u32>, Vec<u32>)> }
impl SceneBuilder {
    pub fn leaf(&mut self, ty: u32, material: u32, p: &[f64]) -> u32 {
        let mut q = [0.0; 12]; q[..p.len()].copy_from_slice(p);
        self.nodes.push(RtbNode { ty, material, first_child: 0, n_children: 0, p: q }); (self.nodes.len() - 1) as u32 }
    pub fn inner(&mut self, ty: u32, material: u32, p: &[f64], kids: &[u32]) -> u32 {
        let first = self.children.len() as u32; self.children.extend_from_slice(kids);
        let mut q = [0.0; 12]; q[..p.len()].copy_from_slice(p);
        self.nodes.push(RtbNode { ty, material, first_child: first, n_children: kids.len() as u32, p: q }); (self.nodes.len() - 1) as u32 }
}

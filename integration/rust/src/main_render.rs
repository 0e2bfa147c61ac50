// raytracer/src/main_render.rs — replaces the render loop of main() (raytracer/src/main.rs:720-784): everything above
// it (scene construction :668-686, camera constants :688-718, image constants :659-663) stays as it is.  The 18-thread
// fan-out per pixel (:730-778) and the mpsc reduction (:767-781) become ONE call; the JPEG writer (:785-798) is unchanged.
// NOT compiled in the build image (no rustc / cargo).
use crate::flatten_impls::{light_records, FlattenHittable};
use crate::gpu::{Gpu, RtbCamera, RtbParams, SceneBuilder};
use crate::hittable_list::HittableList;

pub struct Frame {
    pub width: u32,
    pub height: u32,
    pub samples_per_pixel: u32,
    pub max_depth: i32,
    pub background: [f64; 3],
}

/// `world` = cornell_box() (main.rs:337-433) or any other scene function; `lights` = the proxy list of main.rs:669-686;
/// `devices` = [0] for one B200, [0, 1, .., 7] for the whole box (samples split, one NCCL reduce inside the library).
/// Returns RGB8 rows from the top — the layout `img.get_pixel_mut(i, IMAGE_HEIGHT - j - 1)` fills at main.rs:733,780.
pub fn render_gpu(
    world: &HittableList,
    lights: (&[&crate::aarect::XzRect], &[&crate::sphere::Sphere]),
    cam: RtbCamera,
    frame: &Frame,
    devices: &[i32],
) -> Vec<u8> {
    let mut b = SceneBuilder::default();
    let root = world.flatten(&mut b);
    let light_recs = light_records(lights.0, lights.1);
    let gpu = Gpu::new(devices);
    let prm = RtbParams {
        width: frame.width,
        height: frame.height,
        spp: frame.samples_per_pixel,
        sample_offset: 0,
        total_spp: frame.samples_per_pixel,
        max_depth: frame.max_depth,
        rr_start_depth: 0, // Russian roulette off = the reference's behaviour
        seed: 1,
        background: [frame.background[0] as f32, frame.background[1] as f32, frame.background[2] as f32],
        pool_paths: 0,
        flags: 0,
    };
    let (rgb, stats) = gpu.render(&b, root, &light_recs, &cam, &prm);
    eprintln!(
        "rtb200: {} paths, {} segments in {:.1} ms on {} GPU(s) ({:.0} Mrays/s; NCCL reduce {:.2} ms)",
        stats.paths,
        stats.segments,
        stats.ms_total,
        stats.n_devices,
        stats.segments as f64 / stats.ms_total / 1e3,
        stats.ms_nccl
    );
    rgb
}

// In main(), after `let cam = Camera::new(...)` (main.rs:711-718):
//
//     let frame = Frame { width: IMAGE_WIDTH, height: IMAGE_HEIGHT, samples_per_pixel: SAMPLES_PER_PIXEL, max_depth: MAX_DEPTH,
//                         background: [background.x(), background.y(), background.z()] };
//     let rtb_cam = RtbCamera { lookfrom: [lookfrom.x(), lookfrom.y(), lookfrom.z()], lookat: [lookat.x(), lookat.y(), lookat.z()],
//                               vup: [vup.x(), vup.y(), vup.z()], vfov_deg: vfov, aspect_ratio: ASPECT_RATIO, aperture,
//                               focus_dist: dist_to_focus, time0: 0.0, time1: 1.0 };
//     let rgb = render_gpu(&world, (&[&light_rect], &[&glass_sphere]), rtb_cam, &frame, &[0]);
//     let img: RgbImage = ImageBuffer::from_raw(IMAGE_WIDTH, IMAGE_HEIGHT, rgb).unwrap();
//
// and the existing JPEG encode (main.rs:791-796) runs on `img` unchanged.

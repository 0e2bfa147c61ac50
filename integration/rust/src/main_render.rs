// Replacement for the render loop raytracer/src/main.rs:720-799.  See INTEGRATION.md.
let mut b = gpu::SceneBuilder::default();
let root = world.flatten(&mut b);
let lights = [RtbLight { ty: 0, _pad: 0, p: [213.0, 343.0, 227.0, 332.0, 554.0] },       // main.rs:670-679
              RtbLight { ty: 1, _pad: 0, p: [190.0, 90.0, 190.0, 90.0, 0.0] }];          // main.rs:680-684
unsafe {
    let (mut ctx, mut sc) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(rtb_context_create(0, &mut ctx)); check(rtb_scene_create(ctx, &mut sc));
    check(rtb_scene_set_materials(sc, b.materials.as_ptr(), b.materials.len() as u32));
    check(rtb_scene_set_textures(sc, b.textures.as_ptr(), b.textures.len() as u32));
    check(rtb_scene_set_lights(sc, lights.as_ptr(), 2));
    check(rtb_scene_set_graph(sc, b.nodes.as_ptr(), b.nodes.len() as u32, b.children.as_ptr(), b.children.len() as u32, root));
    check(rtb_scene_commit(sc));
    let cam = RtbCamera { lookfrom: [278.0, 278.0, -800.0], lookat: [278.0, 278.0, 0.0], vup: [0.0, 1.0, 0.0],
                          vfov_deg: 40.0, aspect_ratio: 1.0, aperture: 0.0, focus_dist: 10.0, time0: 0.0, time1: 1.0 };
    let prm = RtbParams { width: IMAGE_WIDTH, height: IMAGE_HEIGHT, spp: SAMPLES_PER_PIXEL, sample_offset: 0,
                          total_spp: SAMPLES_PER_PIXEL, max_depth: MAX_DEPTH, rr_start_depth: 0, seed: 1,
                          background: [0.0; 3], pool_paths: 0, flags: 0 };
    let mut stats = RtbStats::default();
    check(rtb_render(ctx, sc, &cam, &prm, std::ptr::null_mut(), &mut stats));
    let mut rgb = vec![0u8; (IMAGE_WIDTH * IMAGE_HEIGHT * 3) as usize];     // row 0 = top, like img.get_pixel_mut(i, H-1-j)
    check(rtb_finalize_rgb8(ctx, std::ptr::null(), IMAGE_WIDTH, IMAGE_HEIGHT, SAMPLES_PER_PIXEL, rgb.as_mut_ptr()));
    // -> image::RgbImage::from_raw(IMAGE_WIDTH, IMAGE_HEIGHT, rgb) -> JPEG exactly as main.rs:791-796
}
fn check(rc: c_int) { if rc != 0 { panic!("rtb200: {}", unsafe { CStr::from_ptr(rtb_last_error()) }.to_string_lossy()) } }

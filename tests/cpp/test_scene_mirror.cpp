// C++ host-side test: the reference's cornell_box() (raytracer/src/main.rs:337-433) and light list (:669-686) written
// with the C++ mirror of its constructors, handed to librtb200 through the C ABI (host-only scene: no GPU needed here).
// Prints the flattened scene summary and the serialized records so the pytest wrapper can compare them with the Python
// mirror's.  With a GPU (argv[1] == "render") it also renders 16 spp and prints the segment count.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/rtb200_scene.hpp"

using namespace rtb200;

static std::shared_ptr<HittableList> cornell_box() {
  auto objects = HittableList::new_();
  auto red = Lambertian::construct({0.65, 0.05, 0.05});
  auto white = Lambertian::construct({0.73, 0.73, 0.73});
  auto green = Lambertian::construct({0.12, 0.45, 0.15});
  auto light = DiffuseLight::construct_color({15.0, 15.0, 15.0});
  objects->add(YzRect::construct(0.0, 555.0, 0.0, 555.0, 555.0, green));
  objects->add(YzRect::construct(0.0, 555.0, 0.0, 555.0, 0.0, red));
  objects->add(FlipFace::construct(XzRect::construct(213.0, 343.0, 227.0, 332.0, 554.0, light)));
  objects->add(XzRect::construct(0.0, 555.0, 0.0, 555.0, 0.0, white));
  objects->add(XzRect::construct(0.0, 555.0, 0.0, 555.0, 555.0, white));
  objects->add(XyRect::construct(0.0, 555.0, 0.0, 555.0, 555.0, white));
  HittablePtr box1 = Box::construct({0.0, 0.0, 0.0}, {165.0, 330.0, 165.0}, white);
  box1 = RotateY::construct(box1, 15.0);
  box1 = Translate::construct(box1, {265.0, 0.0, 295.0});
  objects->add(box1);
  objects->add(Sphere::construct({190.0, 90.0, 190.0}, 90.0, Dielectric::construct(1.5)));
  return objects;
}

// final_scene() (raytracer/src/main.rs:521-649, commented in the reference) with the C++ mirror.  The random numbers of the
// scene (box heights, the 1000 sphere centres, the Perlin tables) and the earthmap texels come from a file the pytest
// wrapper writes from the Python mirror's scene, so both sides construct the SAME scene and must emit the same records.
struct FinalSceneData {
  std::vector<double> boxes;    // 400 x (x0 y0 z0 x1 y1 z1)
  std::vector<double> centres;  // 1000 x 3
  std::vector<double> ranvec;   // 256 x 3
  std::vector<uint32_t> perm;   // 3 x 256
  std::vector<uint8_t> earth;   // 512 x 1024 x 3
};
static bool read_final_scene_data(const char* path, FinalSceneData& d) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  d.boxes.resize(2400); d.centres.resize(3000); d.ranvec.resize(768); d.perm.resize(768); d.earth.resize(512 * 1024 * 3);
  bool ok = std::fread(d.boxes.data(), 8, 2400, f) == 2400 && std::fread(d.centres.data(), 8, 3000, f) == 3000 &&
            std::fread(d.ranvec.data(), 8, 768, f) == 768 && std::fread(d.perm.data(), 4, 768, f) == 768 &&
            std::fread(d.earth.data(), 1, d.earth.size(), f) == d.earth.size();
  std::fclose(f);
  return ok;
}
static std::shared_ptr<HittableList> final_scene(const FinalSceneData& d) {
  auto boxes1 = HittableList::new_();
  auto ground = Lambertian::construct({0.48, 0.83, 0.53});
  for (int k = 0; k < 400; ++k) {
    const double* b = &d.boxes[6 * k];
    boxes1->add(Box::construct({b[0], b[1], b[2]}, {b[3], b[4], b[5]}, ground));
  }
  auto objects = HittableList::new_();
  objects->add(BVHNode::construct2(boxes1, 0.0, 1.0));
  auto light = DiffuseLight::construct_color({7.0, 7.0, 7.0});
  objects->add(FlipFace::construct(XzRect::construct(123.0, 423.0, 147.0, 412.0, 554.0, light)));
  objects->add(MovingSphere::construct({400.0, 400.0, 200.0}, {430.0, 400.0, 200.0}, 0.0, 1.0, 50.0, Lambertian::construct({0.7, 0.3, 0.1})));
  objects->add(Sphere::construct({260.0, 150.0, 45.0}, 50.0, Dielectric::construct(1.5)));
  objects->add(Sphere::construct({0.0, 150.0, 145.0}, 50.0, Metal::construct({0.8, 0.8, 0.9}, 1.0)));
  HittablePtr boundary = Sphere::construct({360.0, 150.0, 145.0}, 70.0, Dielectric::construct(1.5));
  objects->add(boundary);
  objects->add(ConstantMedium::construct_color(boundary, 0.2, {0.2, 0.4, 0.9}));
  HittablePtr boundary2 = Sphere::construct({0.0, 0.0, 0.0}, 5000.0, Dielectric::construct(1.5));
  objects->add(ConstantMedium::construct_color(boundary2, 0.0001, {1.0, 1.0, 1.0}));
  objects->add(Sphere::construct({400.0, 200.0, 400.0}, 100.0, Lambertian::construct_texture(ImageTexture::construct(d.earth, 1024, 512))));
  auto pertext = std::make_shared<NoiseTexture>();
  pertext->scale = 0.1;
  pertext->ranvec = d.ranvec;
  pertext->perm_x.assign(d.perm.begin(), d.perm.begin() + 256);
  pertext->perm_y.assign(d.perm.begin() + 256, d.perm.begin() + 512);
  pertext->perm_z.assign(d.perm.begin() + 512, d.perm.end());
  objects->add(Sphere::construct({220.0, 280.0, 300.0}, 80.0, Lambertian::construct_texture(pertext)));
  auto boxes2 = HittableList::new_();
  auto white = Lambertian::construct({0.73, 0.73, 0.73});
  for (int k = 0; k < 1000; ++k) boxes2->add(Sphere::construct({d.centres[3 * k], d.centres[3 * k + 1], d.centres[3 * k + 2]}, 10.0, white));
  objects->add(Translate::construct(RotateY::construct(BVHNode::construct2(boxes2, 0.0, 1.0), 15.0), {-100.0, 270.0, 395.0}));
  return objects;
}

static void print_records(const SceneRecords& rec) {
  std::printf("records %zu children %zu root %u\n", rec.nodes.size(), rec.child_index.size(), rec.root);
  for (const rtb_node& n : rec.nodes) {
    std::printf("node %u %u %u %u", n.type, n.material, n.first_child, n.n_children);
    for (double v : n.p) std::printf(" %.17g", v);
    std::printf("\n");
  }
  for (const rtb_material& m : rec.materials) std::printf("material %u %u %.17g\n", m.type, m.texture, m.param);
  for (const rtb_texture& t : rec.textures)
    std::printf("texture %u %u %u %u %.17g %.17g %.17g %.17g\n", t.type, t.even, t.odd, t.table, t.rgb[0], t.rgb[1], t.rgb[2], t.scale);
}

int main(int argc, char** argv) {
  // "final <data file> [render]": the book-2 final scene through the mirror (records + optional 8-spp GPU render)
  if (argc > 2 && !std::strcmp(argv[1], "final")) {
    FinalSceneData d;
    if (!read_final_scene_data(argv[2], d)) { std::printf("cannot read %s\n", argv[2]); return 2; }
    auto fworld = final_scene(d);
    auto flights = HittableList::new_();
    flights->add(XzRect::construct(123.0, 423.0, 147.0, 412.0, 554.0, DiffuseLight::construct_color({7.0, 7.0, 7.0})));
    SceneRecords frec;
    frec.set_world(fworld, flights);
    print_records(frec);
    const bool frender = argc > 3 && !std::strcmp(argv[3], "render");
    rtb_context* fctx = nullptr;
    if (frender && rtb_context_create(0, &fctx) != 0) { std::printf("context: %s\n", rtb_last_error()); return 2; }
    rtb_scene* fscene = nullptr;
    if (rtb_scene_create(fctx, &fscene) != 0) return 3;
    frec.upload(fscene);
    if (rtb_scene_build_bvh(fscene) != 0) { std::printf("build: %s\n", rtb_last_error()); return 4; }
    rtb_scene_info finfo;
    rtb_scene_get_info(fscene, &finfo);
    std::printf("final quads %u spheres %u moving %u media %u prims %u lights %u\n", finfo.n_quads, finfo.n_spheres, finfo.n_moving,
                finfo.n_media, finfo.n_prims, finfo.n_lights);
    if (frender) {
      if (rtb_scene_commit(fscene) != 0) { std::printf("commit: %s\n", rtb_last_error()); return 5; }
      rtb_camera cam{{478, 278, -600}, {278, 278, 0}, {0, 1, 0}, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0};
      rtb_params prm{};
      prm.width = 96; prm.height = 96; prm.spp = 8; prm.total_spp = 8; prm.max_depth = 50; prm.seed = 1;
      rtb_stats st;
      std::vector<float> acc((size_t)prm.width * prm.height * 4);
      if (rtb_render(fctx, fscene, &cam, &prm, acc.data(), &st) != 0) { std::printf("render: %s\n", rtb_last_error()); return 6; }
      double sum = 0;
      for (float v : acc) sum += v;
      std::printf("rendered paths %llu segments %llu sum %.9g\n", (unsigned long long)st.paths, (unsigned long long)st.segments, sum);
    }
    rtb_scene_destroy(fscene);
    if (fctx) rtb_context_destroy(fctx);
    return 0;
  }
  auto world = cornell_box();
  auto lights = HittableList::new_();
  lights->add(XzRect::construct(213.0, 343.0, 227.0, 332.0, 554.0, DiffuseLight::construct_color({15.0, 15.0, 15.0})));
  lights->add(Sphere::construct({190.0, 90.0, 190.0}, 90.0, Dielectric::construct(1.5)));
  SceneRecords rec;
  rec.set_world(world, lights);

  // "multi N": SURVEY §4 tier D1 through the C ABI — one process, N GPUs (rtb_context_create_multi -> ncclCommInitAll, samples
  // split inside rtb_render, one ncclReduce) against the same render on one GPU: same sample set, sums equal to f32
  // summation order
  if (argc > 2 && !std::strcmp(argv[1], "multi")) {
    const int n = std::atoi(argv[2]);
    std::vector<int> ids;
    for (int k = 0; k < n; ++k) ids.push_back(k);
    rtb_context *one = nullptr, *many = nullptr;
    if (rtb_context_create(0, &one) != 0) { std::printf("context: %s\n", rtb_last_error()); return 2; }
    if (rtb_context_create_multi(ids.data(), n, &many) != 0) { std::printf("multi context: %s\n", rtb_last_error()); return 2; }
    rtb_camera cam{{278, 278, -800}, {278, 278, 0}, {0, 1, 0}, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0};
    rtb_params prm{};
    prm.width = 128; prm.height = 96; prm.spp = 50; prm.total_spp = 50; prm.max_depth = 50; prm.seed = 3;  // 50 % n != 0 for n = 4, 8
    std::vector<float> a((size_t)prm.width * prm.height * 4), b(a.size());
    rtb_stats sa, sb;
    rtb_context* ctxs[2] = {one, many};
    std::vector<float>* outs[2] = {&a, &b};
    rtb_stats* sts[2] = {&sa, &sb};
    for (int q = 0; q < 2; ++q) {
      rtb_scene* sc = nullptr;
      if (rtb_scene_create(ctxs[q], &sc) != 0) return 3;
      rec.upload(sc);
      if (rtb_scene_commit(sc) != 0) { std::printf("commit: %s\n", rtb_last_error()); return 5; }
      if (rtb_render(ctxs[q], sc, &cam, &prm, outs[q]->data(), sts[q]) != 0) { std::printf("render: %s\n", rtb_last_error()); return 6; }
      rtb_scene_destroy(sc);
    }
    double sum_a = 0, sum_b = 0, worst = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      sum_a += a[i]; sum_b += b[i];
      const double d = std::fabs((double)a[i] - (double)b[i]) / (std::fabs((double)a[i]) + 1.0);
      if (d > worst) worst = d;
    }
    std::printf("multi devices %u (count %d) segments %llu vs %llu paths %llu vs %llu rel_sum_diff %.3e worst_pixel %.3e ms_nccl %.3f ms_render %.3f\n",
                sb.n_devices, rtb_context_device_count(many), (unsigned long long)sb.segments, (unsigned long long)sa.segments,
                (unsigned long long)sb.paths, (unsigned long long)sa.paths, std::fabs(sum_a - sum_b) / sum_a, worst, sb.ms_nccl, sb.ms_render);
    rtb_context_destroy(many);
    rtb_context_destroy(one);
    return 0;
  }
  const bool render = argc > 1 && !std::strcmp(argv[1], "render");
  rtb_context* ctx = nullptr;
  if (render && rtb_context_create(0, &ctx) != 0) { std::printf("context: %s\n", rtb_last_error()); return 2; }
  rtb_scene* scene = nullptr;
  if (rtb_scene_create(ctx, &scene) != 0) return 3;
  rec.upload(scene);
  if (rtb_scene_build_bvh(scene) != 0) { std::printf("build: %s\n", rtb_last_error()); return 4; }
  rtb_scene_info info;
  rtb_scene_get_info(scene, &info);
  std::printf("quads %u spheres %u prims %u lights %u materials %u textures %u nodes %u\n", info.n_quads, info.n_spheres,
              info.n_prims, info.n_lights, info.n_materials, info.n_textures, info.n_bvh_nodes);
  std::printf("records %zu children %zu root %u\n", rec.nodes.size(), rec.child_index.size(), rec.root);
  for (const rtb_node& n : rec.nodes) {
    std::printf("node %u %u %u %u", n.type, n.material, n.first_child, n.n_children);
    for (double v : n.p) std::printf(" %.17g", v);
    std::printf("\n");
  }
  // error behaviour: a host-only scene cannot be committed
  if (!ctx) {
    int rc = rtb_scene_commit(scene);
    std::printf("commit_without_context %d\n", rc);
  } else {
    if (rtb_scene_commit(scene) != 0) { std::printf("commit: %s\n", rtb_last_error()); return 5; }
    rtb_camera cam{{278, 278, -800}, {278, 278, 0}, {0, 1, 0}, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0};
    rtb_params prm{};
    prm.width = 64; prm.height = 64; prm.spp = 16; prm.total_spp = 16; prm.max_depth = 50; prm.seed = 1;
    rtb_stats st;
    if (rtb_render(ctx, scene, &cam, &prm, nullptr, &st) != 0) { std::printf("render: %s\n", rtb_last_error()); return 6; }
    std::printf("rendered paths %llu segments %llu\n", (unsigned long long)st.paths, (unsigned long long)st.segments);
  }
  rtb_scene_destroy(scene);
  if (ctx) rtb_context_destroy(ctx);
  return 0;
}

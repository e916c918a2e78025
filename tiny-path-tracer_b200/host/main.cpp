// host/main.cpp -- the driver, same flow as the reference's main.cpp:24-247:
//   read ./config.ini (same keys, same defaults) -> echo it -> build the scene with the host
//   classes -> camera -> light-sampling list -> [flatten + libtpt.so render on the GPU(s)]
//   -> img.ppm (+ img_k.ppm bonus pictures) -> "time: X s" -> convert ... +append img.jpg
// Only the bracketed step differs: it replaces the row-per-thread sample loop
// (main.cpp:109-175) with the sm_100a core behind include/tpt.h.
//
// Additional, optional keys (absent => reference behaviour):
//   [SCENE] name = cornell_box | sphere_cornell_box | random_scene | random_scene_list |
//                  two_perlin_spheres | two_checker_spheres | light_spheres | earth
//                  | cornell_box_smoke | oneweek_final
//           background = black | sky        image = earthmap.jpg
//           lights = reference | auto       (auto: the scene's own diffuse_light xz_rects / spheres
//                                            instead of the hard-coded list of main.cpp:99-106)
//   [CAMERA] lookfrom = x,y,z   lookat = x,y,z   vup = x,y,z   focus_dist = <f>
//   [GPU]   mode = fast | parity    kernel = mega | wavefront   seed = <u64>   gpus = <n>
//   [OUTPUT] jpeg = native | convert   quality = 92   ppm = p3 | p6
#include "tpt.h"
#include "tpt_flatten.h"
#include "tpt_image_io.h"
#include "inipp.h" // third_party/inipp.h, vendored verbatim (MIT)
#include "tpt_scene.h"

#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

int main(int, char **) {
  // the CUDA context (a few hundred ms in a fresh process) comes up on a second host thread while this one reads
  // config.ini and builds the scene; joined before the first upload
  std::thread gpu_warm([] { tpt_device_warm(0); });
  struct join_on_exit {
    std::thread &t;
    ~join_on_exit() {
      if (t.joinable()) t.join();
    }
  } gpu_warm_guard{gpu_warm};
  std::vector<std::string> filenames;
  filenames.push_back("img.ppm");

  std::ifstream config("config.ini");
  // defaults of main.cpp:31-40
  int nx = 400, ny = 200, ns = 10;
  float aperture = 0.2f, time0 = 0.0f, time1 = 1.0f;
  int sample_max_recurse_depth = 50;
  float fov = 90.0f;
  int allow_bonus_pic = 0, bonus_pic = 10;
  std::string scene_name = "cornell_box", background = "black", image_file = "earthmap.jpg";
  std::string mode = "fast", kernel = "wavefront";
  std::string lights = "reference";                  // reference: main.cpp:99-106 ; auto: the scene's own lamps
  std::string jpeg_tool = "native", ppm_format = "p3"; // [OUTPUT]
  std::string cam_lookfrom, cam_lookat, cam_vup;     // [CAMERA] "x,y,z"; empty: the scene's camera of main.cpp:85-91
  float cam_focus = 0.0f;
  int jpeg_quality = 92;
  unsigned long long seed = 0x5EEDULL;
  int gpus = 1;

  inipp::Ini<char> ini;
  ini.parse(config);
  inipp::extract(ini.sections["DEFAULT"]["width"], nx);
  inipp::extract(ini.sections["DEFAULT"]["height"], ny);
  inipp::extract(ini.sections["DEFAULT"]["sample"], ns);
  inipp::extract(ini.sections["DEFAULT"]["recur_depth"], sample_max_recurse_depth);
  inipp::extract(ini.sections["DEFAULT"]["fov"], fov);
  inipp::extract(ini.sections["DEFAULT"]["bonus_pic"], bonus_pic);
  inipp::extract(ini.sections["DEFAULT"]["allow_bonus_pic"], allow_bonus_pic);
  inipp::extract(ini.sections["BLUR"]["aperture"], aperture);
  inipp::extract(ini.sections["CAM_MOTION"]["start_time"], time0);
  inipp::extract(ini.sections["CAM_MOTION"]["end_time"], time1);
  // optional keys: read only when present (the string form of extract() would otherwise replace a
  // default with the empty value operator[] inserts)
  auto opt = [&ini](const char *section, const char *key, auto &dst) {
    auto sec = ini.sections.find(section);
    if (sec == ini.sections.end()) return;
    auto kv = sec->second.find(key);
    if (kv != sec->second.end()) inipp::extract(kv->second, dst);
  };
  opt("SCENE", "name", scene_name);
  opt("SCENE", "background", background);
  opt("SCENE", "image", image_file);
  opt("SCENE", "lights", lights);
  opt("CAMERA", "lookfrom", cam_lookfrom);
  opt("CAMERA", "lookat", cam_lookat);
  opt("CAMERA", "vup", cam_vup);
  opt("CAMERA", "focus_dist", cam_focus);
  opt("GPU", "mode", mode);
  opt("GPU", "kernel", kernel);
  opt("GPU", "seed", seed);
  opt("GPU", "gpus", gpus);
  opt("OUTPUT", "jpeg", jpeg_tool);
  opt("OUTPUT", "ppm", ppm_format);
  opt("OUTPUT", "quality", jpeg_quality);
  ini.generate(std::cout);

  std::cout << "<===========>" << std::endl;
#ifdef DEBUG_MODE
  std::cout << "IN DEBUG MODE" << std::endl;
#else
  std::cout << "IN RELEASE MODE" << std::endl;
#endif
  std::cout << "Threads num: " << std::thread::hardware_concurrency() << std::endl;
  std::cout << "GPUs: " << tpt_device_count() << " (using " << gpus << ", " << mode << "/" << kernel
            << ")" << std::endl;
  std::cout << "<===========>" << std::endl;

  if (allow_bonus_pic && (bonus_pic <= 0 || ns / bonus_pic <= 0)) {
    // the reference divides by zero here (main.cpp:113,127): refuse instead
    std::cerr << "sample must be >= bonus_pic when allow_bonus_pic is set" << std::endl;
    return 2;
  }

  // ---- scene, camera, light list: host classes, as main.cpp:71-106 ----
  hitable *world = nullptr;
  vec3 lookfrom(0, 0, 800), lookat(0, 0, 0);
  float dist_to_focus = 10.0f;
  if (scene_name == "cornell_box") {
    world = cornell_box();
  } else {
    // the alternatives the reference keeps commented out (main.cpp:71-84)
    if (scene_name == "sphere_cornell_box") world = sphere_cornell_box();
    else if (scene_name == "random_scene") world = random_scene();
    else if (scene_name == "two_perlin_spheres") world = two_perlin_spheres();
    else if (scene_name == "two_checker_spheres") world = two_checker_spheres();
    else if (scene_name == "light_spheres") world = light_spheres();
    else if (scene_name == "cornell_box_smoke") world = cornell_box_smoke();
    else if (scene_name == "oneweek_final") world = oneweek_final();
    else if (scene_name == "earth") {
      int w, h, ch;
      unsigned char *data = load_image_texture(image_file, w, h, ch);
      if (!data) {
        std::cerr << "cannot load " << image_file << std::endl;
        return 2;
      }
      world = new sphere({0, 0, 0}, 3, new lambertian(new image_texture(data, w, h)));
    } else {
      std::cerr << "unknown scene " << scene_name << std::endl;
      return 2;
    }
    if (scene_name == "oneweek_final") { // the book's camera for this scene (not in the reference's main.cpp)
      lookfrom = vec3(478, 278, -600);
      lookat = vec3(278, 278, 0);
      dist_to_focus = 10.0f;
    } else if (scene_name != "sphere_cornell_box" && scene_name != "cornell_box_smoke") {
      lookfrom = vec3(13, 2, 3);
      dist_to_focus = (lookfrom - lookat).length();
    }
  }
  vec3 vup(0, 1, 0);
  auto parse_vec3 = [](const std::string &text, vec3 &v) {
    float x, y, z;
    if (std::sscanf(text.c_str(), " %f , %f , %f", &x, &y, &z) != 3) return false;
    v = vec3(x, y, z);
    return true;
  };
  if ((!cam_lookfrom.empty() && !parse_vec3(cam_lookfrom, lookfrom)) || (!cam_lookat.empty() && !parse_vec3(cam_lookat, lookat)) ||
      (!cam_vup.empty() && !parse_vec3(cam_vup, vup))) {
    std::cerr << "[CAMERA] lookfrom / lookat / vup must be x,y,z" << std::endl;
    return 2;
  }
  if (cam_focus > 0) dist_to_focus = cam_focus;
  else if (!cam_lookfrom.empty() || !cam_lookat.empty()) dist_to_focus = (lookfrom - lookat).length();
  camera cam(lookfrom, lookat, vup, fov, float(nx) / (float)ny, aperture, dist_to_focus,
             time0, time1);

  xz_rect light_shape(-100, 100, -150, -50, 298, nullptr);
  sphere sphere_shape(vec3(120, -50, 40), 120, nullptr);
  hitable *a[2] = {&light_shape, &sphere_shape};
  hitable_list hlist(a, 2);

  gpu_warm.join();
  auto start = std::chrono::high_resolution_clock::now();

  // ---- the replaced block: main.cpp:109-175 ----
  tpt::FlatScene flat;
  std::string err;
  if (!tpt::flatten_scene(world, &hlist, background == "sky" ? TPT_BG_SKY : TPT_BG_BLACK, flat, err)) {
    std::cerr << "flatten: " << err << std::endl;
    return 3;
  }
  if (lights == "auto") {
    std::vector<tpt_light> own = tpt::derive_light_list(flat);
    std::cout << "light list: " << own.size() << " emissive primitive(s) found"
              << (own.empty() ? ", keeping the reference list" : "") << std::endl;
    if (!own.empty()) flat.lights = own;
  }
  tpt_scene_desc desc = flat.desc();
  if (gpus < 1) gpus = 1;
  if (gpus > tpt_device_count()) gpus = tpt_device_count() > 0 ? tpt_device_count() : 1;
  std::vector<tpt_scene *> scenes(gpus, nullptr); // the (tiny) scene is replicated on every GPU used
  for (int g = 0; g < gpus; g++) {
    if (tpt_scene_create(&desc, g, &scenes[g]) != TPT_OK) {
      std::cerr << "tpt_scene_create: " << tpt_last_error() << std::endl;
      return 3;
    }
  }
  tpt_scene *scene = scenes[0];
  tpt_camera cam_desc = tpt::make_camera_desc(cam);
  tpt_render_params rp{};
  rp.nx = nx;
  rp.ny = ny;
  rp.ns = ns;
  rp.max_depth = sample_max_recurse_depth;
  rp.slices = allow_bonus_pic ? bonus_pic : 1;
  rp.mode = mode == "parity" ? TPT_MODE_PARITY : TPT_MODE_FAST;
  rp.kernel = kernel == "wavefront" ? TPT_KERNEL_WAVEFRONT : TPT_KERNEL_MEGA;
  rp.seed_lo = (uint32_t)seed;
  rp.seed_hi = (uint32_t)(seed >> 32);
  rp.t_min = 0.001f;
  rp.part_index = 0;
  rp.part_count = 1;
  rp.device = 0;
  std::vector<uint8_t> rgb8((size_t)nx * ny * 3);
  std::vector<uint8_t> rgb8_slices(allow_bonus_pic ? (size_t)rp.slices * nx * ny * 3 : 0);
  tpt_image img{};
  img.rgb8 = rgb8.data();
  img.rgb8_slices = allow_bonus_pic ? rgb8_slices.data() : nullptr;
  // one GPU: tpt_render; several: static tile split + work stealing + NVLink gather
  int rc = gpus > 1 ? tpt_render_multi(scenes.data(), gpus, &cam_desc, &rp, &img) : tpt_render(scene, &cam_desc, &rp, &img);
  if (rc != TPT_OK) {
    std::cerr << "tpt_render: " << tpt_last_error() << std::endl;
    return 4;
  }
  tpt_stats st{};
  tpt_get_stats(scene, &st);
  std::cout << "gpu render: " << st.render_ms / 1000.0 << " s, " << st.paths / (st.render_ms * 1e3)
            << " Mpaths/s, " << st.rays / (st.render_ms * 1e3) << " Mrays/s" << std::endl;
  auto rendered = std::chrono::high_resolution_clock::now();
  std::cout << "flatten + upload + render + download: "
            << std::chrono::duration_cast<std::chrono::microseconds>(rendered - start).count() / 1e6 << " s" << std::endl;

  // ---- output, as main.cpp:176-245 ----
  const bool p6 = ppm_format == "p6"; // binary PPM instead of the reference's text form
  if (p6) tpt::write_ppm_binary(filenames[0], rgb8.data(), nx, ny);
  else tpt::write_ppm_main(filenames[0], rgb8.data(), nx, ny);
  std::vector<const uint8_t *> pictures{rgb8.data()};
  if (allow_bonus_pic) {
    for (int k = 0; k < bonus_pic; k++) {
      std::string file = "img_" + std::to_string(k) + ".ppm";
      filenames.push_back(file);
      const uint8_t *slice = rgb8_slices.data() + (size_t)k * nx * ny * 3;
      pictures.push_back(slice);
      if (p6) tpt::write_ppm_binary(file, slice, nx, ny);
      else tpt::write_ppm_bonus(file, slice, nx, ny);
    }
  }
  auto end = std::chrono::high_resolution_clock::now();
  std::cout << "output files: " << std::chrono::duration_cast<std::chrono::microseconds>(end - rendered).count() / 1e6 << " s"
            << std::endl;
  std::cout << "time: "

            << std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count() / 1000.0f
            << " s" << std::endl;
  for (tpt_scene *sc : scenes) tpt_scene_destroy(sc);
  // main.cpp:224-245 shells out to `convert ... +append img.jpg`; the built-in writer produces the
  // same side-by-side JPEG from the pixels in memory ([OUTPUT] jpeg=convert keeps the old route)
  std::string stem = filenames[0].substr(0, filenames[0].find_first_of("."));
  if (jpeg_tool == "convert" || !tpt::write_contact_sheet(stem + ".jpg", pictures, nx, ny, jpeg_quality)) {
#ifdef __linux__
    tpt::merge_with_convert(filenames);
#endif
  } else {
    std::cout << "Merging pics:" << "\n" << pictures.size() << " picture(s) +append " << stem << ".jpg (built-in JPEG writer)" << std::endl;
  }
  return 0;
}

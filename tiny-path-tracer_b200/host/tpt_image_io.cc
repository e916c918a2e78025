// host/tpt_image_io.cc -- see tpt_image_io.h
#include "tpt_image_io.h"

#define STB_IMAGE_IMPLEMENTATION // as src/utils.cc:11-12
#include "stb_image.h"
#include "tpt_scene.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

namespace tpt {

namespace {
// fast decimal formatting: the reference's ofstream << int path costs ~4 s at 1200x1200
// (SURVEY 6.2); at sub-second render times the writer must not dominate.
inline char *put_u8(char *p, unsigned v) {
  if (v >= 100) {
    *p++ = char('0' + v / 100);
    v %= 100;
    *p++ = char('0' + v / 10);
    *p++ = char('0' + v % 10);
  } else if (v >= 10) {
    *p++ = char('0' + v / 10);
    *p++ = char('0' + v % 10);
  } else {
    *p++ = char('0' + v);
  }
  return p;
}
bool write_p3(const std::string &path, const uint8_t *rgb8, int nx, int ny, bool pixel_per_line) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "P3\n%d %d\n255\n", nx, ny);
  std::vector<char> line((size_t)nx * 13 + 8);
  for (int j = ny - 1; j >= 0; j--) { // top row first (main.cpp:183)
    char *p = line.data();
    const uint8_t *row = rgb8 + (size_t)j * nx * 3;
    for (int i = 0; i < nx; i++) {
      for (int c = 0; c < 3; c++) {
        p = put_u8(p, row[3 * i + c]);
        *p++ = ' ';
      }
      if (pixel_per_line) *p++ = '\n';
    }
    std::fwrite(line.data(), 1, (size_t)(p - line.data()), f);
  }
  return std::fclose(f) == 0;
}
} // namespace

bool write_ppm_main(const std::string &path, const uint8_t *rgb8, int nx, int ny) {
  return write_p3(path, rgb8, nx, ny, true);
}
bool write_ppm_bonus(const std::string &path, const uint8_t *rgb8, int nx, int ny) {
  return write_p3(path, rgb8, nx, ny, false);
}

int merge_with_convert(const std::vector<std::string> &files) {
  if (files.empty()) return -1;
  std::string list;
  for (const std::string &f : files) list += f + " ";
  std::string stem = files[0].substr(0, files[0].find_first_of("."));
  std::string command = "convert " + list + " +append " + stem + ".jpg";
  std::cout << "Merging pics:" << "\n" << command << std::endl;
  int ret = std::system(command.c_str());
  if (ret == -1) perror("os.system error");
  return ret;
}

bool read_ppm(const std::string &path, std::vector<uint8_t> &rgb, int &w, int &h) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return false;
  std::string magic;
  in >> magic;
  if (magic != "P6" && magic != "P3") return false;
  auto next_int = [&](int &v) {
    for (;;) {
      in >> std::ws;
      if (in.peek() == '#') {
        std::string skip;
        std::getline(in, skip);
      } else
        break;
    }
    return bool(in >> v);
  };
  int maxv = 0;
  if (!next_int(w) || !next_int(h) || !next_int(maxv) || w <= 0 || h <= 0 || maxv != 255) return false;
  rgb.resize((size_t)w * h * 3);
  if (magic == "P6") {
    in.get(); // single whitespace after maxval
    in.read(reinterpret_cast<char *>(rgb.data()), (std::streamsize)rgb.size());
    return bool(in);
  }
  for (size_t i = 0; i < rgb.size(); i++) {
    int v;
    if (!(in >> v)) return false;
    rgb[i] = (uint8_t)v;
  }
  return true;
}

} // namespace tpt

// Texture input, exactly as the reference: stb_image (third_party/stb_image.h, vendored verbatim,
// v2.23, public domain) behind the same wrapper, src/utils.cc:236-240. Returns stbi's malloc'ed
// memory (channels as in the file; the scene builders expect 3, src/texture.cc:27-42), or nullptr.
unsigned char *load_image_texture(std::string filename, int &width, int &height, int &channels) {
  return stbi_load(filename.c_str(), &width, &height, &channels, 0);
}

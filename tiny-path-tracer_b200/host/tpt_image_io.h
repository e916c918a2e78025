// host/tpt_image_io.h -- picture output in the reference's formats + texture input.
#ifndef TPT_HOST_IMAGE_IO_H_
#define TPT_HOST_IMAGE_IO_H_
#include <cstdint>
#include <string>
#include <vector>

namespace tpt {
// Main picture: "P3\n<nx> <ny>\n255\n" then ONE PIXEL PER LINE "r g b \n", rows top to bottom
// (main.cpp:69,183-189). rgb8 is [ny][nx][3] with row 0 = bottom row.
bool write_ppm_main(const std::string &path, const uint8_t *rgb8, int nx, int ny);
// Bonus pictures: same header, all pixels on one long line "r g b " (main.cpp:197-211).
bool write_ppm_bonus(const std::string &path, const uint8_t *rgb8, int nx, int ny);
// `convert a.ppm b.ppm ... +append img.jpg` (main.cpp:224-245); returns the system() status.
int merge_with_convert(const std::vector<std::string> &files);
// The same output stage without ImageMagick (tpt_jpeg_enc.cc): the pictures side by side, left to
// right, as one baseline JPEG. Buffers are the library's rgb8 layout ([ny][nx][3], row 0 = bottom).
bool write_contact_sheet(const std::string &path, const std::vector<const uint8_t *> &rgb8_bottom_up, int nx, int ny,
                         int quality = 92);
// baseline JPEG writer, rows top to bottom; quality follows the IJG scale (convert's default is 92)
bool encode_jpeg(const uint8_t *rgb, int w, int h, int quality, std::vector<uint8_t> &out);
bool write_jpeg(const std::string &path, const uint8_t *rgb, int w, int h, int quality = 92);
void append_pictures(const std::vector<const uint8_t *> &pictures, const std::vector<int> &widths,
                     const std::vector<int> &heights, std::vector<uint8_t> &out, int &w, int &h);
// binary PPM ("P6"), rows top to bottom; rgb8 is the library's bottom-up layout
bool write_ppm_binary(const std::string &path, const uint8_t *rgb8, int nx, int ny);
// PPM (P6/P3) reader used by tests and as a texture source
bool read_ppm(const std::string &path, std::vector<uint8_t> &rgb, int &w, int &h);
} // namespace tpt
#endif

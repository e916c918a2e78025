// host/tpt_jpeg_enc.cc -- baseline JPEG writer for the output stage.
//
// The reference ends by shelling out to ImageMagick: `convert img.ppm img_0.ppm ... +append
// img.jpg` (main.cpp:224-245), i.e. it re-reads its own P3 text files, puts the pictures side by
// side and encodes a JPEG. SURVEY 8f(3) asks for that stage without the external tool: at
// sub-second render times the text round trip and the process spawn dominate. This file encodes
// the side-by-side picture straight from the 8-bit pixels the library returned.
//
// Format: ITU T.81 baseline sequential DCT, 8-bit, three components without chroma subsampling
// (what `convert` writes for a PPM source at its default quality 92), JFIF 1.01 header. The
// quantisation tables are the T.81 annex K examples scaled by the IJG quality rule; the Huffman
// tables are built per picture from the symbol statistics with the annex K.2 procedure (two
// passes over the coefficients), so no fixed code tables are embedded. tests/test_host_frontend.py
// decodes the result with PIL and with the vendored stb_image and checks the PSNR.
#include "tpt_image_io.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

namespace tpt {

namespace {

// T.81 annex K.1 example tables (natural order)
const uint8_t kLumaQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                            14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                            18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                            49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                              24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                              99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                              99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

void zigzag_order(uint8_t zz[64]) { // index in natural order of the k-th coefficient of the scan
  int r = 0, c = 0;
  for (int k = 0; k < 64; k++) {
    zz[k] = (uint8_t)(r * 8 + c);
    if ((r + c) % 2 == 0) { // moving up-right
      if (c == 7) r++;
      else if (r == 0) c++;
      else { r--; c++; }
    } else { // moving down-left
      if (r == 7) c++;
      else if (c == 0) r++;
      else { r++; c--; }
    }
  }
}

void scaled_table(const uint8_t base[64], int quality, uint8_t out[64]) {
  quality = std::min(100, std::max(1, quality));
  const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
  for (int i = 0; i < 64; i++) out[i] = (uint8_t)std::min(255, std::max(1, (base[i] * scale + 50) / 100));
}

// forward 8x8 DCT-II (orthonormal form of T.81 A.3.3), separable, float
struct Dct {
  float c[8][8];
  Dct() {
    for (int u = 0; u < 8; u++)
      for (int x = 0; x < 8; x++)
        c[u][x] = (u == 0 ? std::sqrt(0.125f) : 0.5f) * std::cos((2 * x + 1) * u * 3.14159265358979323846f / 16.0f);
  }
  void forward(const float in[64], float out[64]) const {
    float tmp[64];
    for (int y = 0; y < 8; y++)
      for (int u = 0; u < 8; u++) {
        float s = 0;
        for (int x = 0; x < 8; x++) s += c[u][x] * in[y * 8 + x];
        tmp[y * 8 + u] = s;
      }
    for (int u = 0; u < 8; u++)
      for (int v = 0; v < 8; v++) {
        float s = 0;
        for (int y = 0; y < 8; y++) s += c[v][y] * tmp[y * 8 + u];
        out[v * 8 + u] = s;
      }
  }
};

inline int bit_size(int v) { // category of T.81 F.1.2.1: bits needed for |v|
  v = v < 0 ? -v : v;
  int n = 0;
  while (v) {
    n++;
    v >>= 1;
  }
  return n;
}

// Huffman code lengths from symbol frequencies, limited to 16 bits (T.81 annex K.2, figures
// K.1-K.4). Symbol 256 is the reserved code point that keeps the all-ones code unused.
struct HuffTable {
  uint8_t counts[16]; // codes per length 1..16
  uint8_t symbols[256];
  int n_symbols = 0;
  uint16_t code[256];
  uint8_t length[256];

  void from_frequencies(const long freq_in[256]) {
    long freq[257];
    int codesize[257], others[257];
    for (int i = 0; i < 256; i++) freq[i] = freq_in[i];
    freq[256] = 1;
    for (int i = 0; i < 257; i++) {
      codesize[i] = 0;
      others[i] = -1;
    }
    for (;;) {
      int c1 = -1, c2 = -1;
      long v = 0;
      for (int i = 0; i <= 256; i++) // least frequent, larger index on ties
        if (freq[i] && (c1 < 0 || freq[i] <= v)) {
          v = freq[i];
          c1 = i;
        }
      v = 0;
      for (int i = 0; i <= 256; i++)
        if (freq[i] && i != c1 && (c2 < 0 || freq[i] <= v)) {
          v = freq[i];
          c2 = i;
        }
      if (c2 < 0) break;
      freq[c1] += freq[c2];
      freq[c2] = 0;
      for (codesize[c1]++; others[c1] >= 0;) {
        c1 = others[c1];
        codesize[c1]++;
      }
      others[c1] = c2;
      for (codesize[c2]++; others[c2] >= 0;) {
        c2 = others[c2];
        codesize[c2]++;
      }
    }
    int bits[66] = {0};
    for (int i = 0; i <= 256; i++)
      if (codesize[i]) bits[std::min(codesize[i], 64)]++;
    for (int i = 64; i > 16; i--) // figure K.3: fold lengths above 16 back
      while (bits[i] > 0) {
        int j = i - 2;
        while (bits[j] == 0) j--;
        bits[i] -= 2;
        bits[i - 1]++;
        bits[j + 1] += 2;
        bits[j]--;
      }
    int top = 16;
    while (bits[top] == 0) top--;
    bits[top]--; // the reserved code point
    for (int i = 0; i < 16; i++) counts[i] = (uint8_t)bits[i + 1];
    n_symbols = 0; // figure K.4: symbols by increasing code length
    for (int len = 1; len <= 64; len++)
      for (int s = 0; s < 256; s++)
        if (codesize[s] == len) symbols[n_symbols++] = (uint8_t)s;
    // canonical codes (annex C)
    std::memset(length, 0, sizeof(length));
    int k = 0;
    unsigned next = 0;
    for (int len = 1; len <= 16; len++) {
      for (int i = 0; i < counts[len - 1]; i++, k++) {
        code[symbols[k]] = (uint16_t)next++;
        length[symbols[k]] = (uint8_t)len;
      }
      next <<= 1;
    }
  }
};

struct BitWriter {
  std::vector<uint8_t> &out;
  uint32_t acc = 0;
  int n = 0;
  explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
  void put(unsigned bits, int count) {
    acc = (acc << count) | (bits & ((1u << count) - 1u));
    n += count;
    while (n >= 8) {
      uint8_t b = (uint8_t)(acc >> (n - 8));
      out.push_back(b);
      if (b == 0xFF) out.push_back(0); // byte stuffing (F.1.2.3)
      n -= 8;
    }
  }
  void flush() {
    if (n > 0) put(0x7F, 8 - n); // pad with ones
  }
};

void put16(std::vector<uint8_t> &o, int v) {
  o.push_back((uint8_t)(v >> 8));
  o.push_back((uint8_t)v);
}

} // namespace

bool encode_jpeg(const uint8_t *rgb, int w, int h, int quality, std::vector<uint8_t> &out) {
  if (!rgb || w <= 0 || h <= 0 || w > 65535 || h > 65535) return false;
  uint8_t zz[64], q[2][64];
  zigzag_order(zz);
  scaled_table(kLumaQ, quality, q[0]);
  scaled_table(kChromaQ, quality, q[1]);
  const int bw = (w + 7) / 8, bh = (h + 7) / 8;
  const size_t n_blocks = (size_t)bw * bh;
  // quantised coefficients in scan order: [block][component][64]
  std::vector<int16_t> coef(n_blocks * 3 * 64);
  const Dct dct;
  auto rows = [&](int by0, int by1) {
    float comp[3][64], freq[64];
    for (int by = by0; by < by1; by++)
      for (int bx = 0; bx < bw; bx++) {
        for (int y = 0; y < 8; y++)
          for (int x = 0; x < 8; x++) {
            const int sx = std::min(w - 1, bx * 8 + x), sy = std::min(h - 1, by * 8 + y); // edge replication
            const uint8_t *p = rgb + ((size_t)sy * w + sx) * 3;
            const float r = p[0], g = p[1], b = p[2];
            // JFIF: full-range BT.601
            comp[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
            comp[1][y * 8 + x] = -0.168735892f * r - 0.331264108f * g + 0.5f * b;
            comp[2][y * 8 + x] = 0.5f * r - 0.418687589f * g - 0.081312411f * b;
          }
        int16_t *dst = &coef[((size_t)by * bw + bx) * 3 * 64];
        for (int c = 0; c < 3; c++) {
          dct.forward(comp[c], freq);
          const uint8_t *qt = q[c ? 1 : 0];
          for (int k = 0; k < 64; k++) dst[c * 64 + k] = (int16_t)std::lrintf(freq[zz[k]] / qt[zz[k]]);
        }
      }
  };
  {
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    nt = (unsigned)std::min<int>((int)nt, bh);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++) pool.emplace_back(rows, (int)((long)bh * t / nt), (int)((long)bh * (t + 1) / nt));
    for (auto &t : pool) t.join();
  }
  // pass 1: symbol statistics (DC differences per component, AC run/size)
  long f_dc[2][256] = {{0}}, f_ac[2][256] = {{0}};
  auto walk = [&](auto &&dc_sym, auto &&ac_sym) {
    int pred[3] = {0, 0, 0};
    for (size_t b = 0; b < n_blocks; b++)
      for (int c = 0; c < 3; c++) {
        const int16_t *v = &coef[(b * 3 + c) * 64];
        const int t = c ? 1 : 0;
        const int diff = v[0] - pred[c];
        pred[c] = v[0];
        dc_sym(t, bit_size(diff), diff);
        int run = 0;
        int last = 63;
        while (last > 0 && v[last] == 0) last--;
        for (int k = 1; k <= last; k++) {
          if (v[k] == 0) {
            run++;
            continue;
          }
          while (run > 15) {
            ac_sym(t, 0xF0, 0, 0); // ZRL
            run -= 16;
          }
          const int s = bit_size(v[k]);
          ac_sym(t, (run << 4) | s, s, v[k]);
          run = 0;
        }
        if (last < 63) ac_sym(t, 0x00, 0, 0); // EOB
      }
  };
  walk([&](int t, int s, int) { f_dc[t][s]++; }, [&](int t, int sym, int, int) { f_ac[t][sym]++; });
  HuffTable hdc[2], hac[2];
  for (int t = 0; t < 2; t++) {
    hdc[t].from_frequencies(f_dc[t]);
    hac[t].from_frequencies(f_ac[t]);
  }
  // headers
  out.clear();
  out.reserve(n_blocks * 24 + 1024);
  const uint8_t soi_app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  out.insert(out.end(), soi_app0, soi_app0 + sizeof(soi_app0));
  for (int t = 0; t < 2; t++) { // DQT, 8-bit entries in scan order
    out.push_back(0xFF);
    out.push_back(0xDB);
    put16(out, 67);
    out.push_back((uint8_t)t);
    for (int k = 0; k < 64; k++) out.push_back(q[t][zz[k]]);
  }
  out.push_back(0xFF); // SOF0
  out.push_back(0xC0);
  put16(out, 17);
  out.push_back(8);
  put16(out, h);
  put16(out, w);
  out.push_back(3);
  for (int c = 0; c < 3; c++) {
    out.push_back((uint8_t)(c + 1));
    out.push_back(0x11);
    out.push_back((uint8_t)(c ? 1 : 0));
  }
  auto dht = [&](int cls, int id, const HuffTable &H) {
    out.push_back(0xFF);
    out.push_back(0xC4);
    put16(out, 19 + H.n_symbols);
    out.push_back((uint8_t)((cls << 4) | id));
    out.insert(out.end(), H.counts, H.counts + 16);
    out.insert(out.end(), H.symbols, H.symbols + H.n_symbols);
  };
  for (int t = 0; t < 2; t++) {
    dht(0, t, hdc[t]);
    dht(1, t, hac[t]);
  }
  out.push_back(0xFF); // SOS
  out.push_back(0xDA);
  put16(out, 12);
  out.push_back(3);
  for (int c = 0; c < 3; c++) {
    out.push_back((uint8_t)(c + 1));
    out.push_back((uint8_t)(c ? 0x11 : 0x00));
  }
  out.push_back(0);
  out.push_back(63);
  out.push_back(0);
  // pass 2: entropy-coded segment
  BitWriter bw_out(out);
  auto magnitude = [](int v, int s) { return (unsigned)(v < 0 ? v + (1 << s) - 1 : v); };
  walk(
      [&](int t, int s, int diff) {
        bw_out.put(hdc[t].code[s], hdc[t].length[s]);
        if (s) bw_out.put(magnitude(diff, s), s);
      },
      [&](int t, int sym, int s, int v) {
        bw_out.put(hac[t].code[sym], hac[t].length[sym]);
        if (s) bw_out.put(magnitude(v, s), s);
      });
  bw_out.flush();
  out.push_back(0xFF);
  out.push_back(0xD9);
  return true;
}

bool write_jpeg(const std::string &path, const uint8_t *rgb, int w, int h, int quality) {
  std::vector<uint8_t> bytes;
  if (!encode_jpeg(rgb, w, h, quality, bytes)) return false;
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
  return std::fclose(f) == 0 && ok;
}

bool write_ppm_binary(const std::string &path, const uint8_t *rgb8, int nx, int ny) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "P6\n%d %d\n255\n", nx, ny);
  bool ok = true;
  for (int j = ny - 1; j >= 0 && ok; j--) // top row first, like the text writers
    ok = std::fwrite(rgb8 + (size_t)j * nx * 3, 1, (size_t)nx * 3, f) == (size_t)nx * 3;
  return std::fclose(f) == 0 && ok;
}

// `+append`: pictures side by side, left to right, top edges aligned; the canvas is as tall as
// the tallest picture and the unused area is white (ImageMagick's default background).
void append_pictures(const std::vector<const uint8_t *> &pictures, const std::vector<int> &widths,
                     const std::vector<int> &heights, std::vector<uint8_t> &out, int &w, int &h) {
  w = 0;
  h = 0;
  for (size_t i = 0; i < pictures.size(); i++) {
    w += widths[i];
    h = std::max(h, heights[i]);
  }
  out.assign((size_t)w * h * 3, 255);
  int x0 = 0;
  for (size_t i = 0; i < pictures.size(); i++) {
    for (int y = 0; y < heights[i]; y++)
      std::memcpy(&out[((size_t)y * w + x0) * 3], pictures[i] + (size_t)y * widths[i] * 3, (size_t)widths[i] * 3);
    x0 += widths[i];
  }
}

} // namespace tpt

namespace tpt {
bool write_contact_sheet(const std::string &path, const std::vector<const uint8_t *> &rgb8_bottom_up, int nx, int ny,
                         int quality) {
  if (rgb8_bottom_up.empty()) return false;
  const int w = nx * (int)rgb8_bottom_up.size();
  std::vector<uint8_t> sheet((size_t)w * ny * 3);
  for (size_t i = 0; i < rgb8_bottom_up.size(); i++)
    for (int y = 0; y < ny; y++) // row 0 of the library's buffers is the bottom row (main.cpp:122,183)
      std::memcpy(&sheet[((size_t)y * w + i * nx) * 3], rgb8_bottom_up[i] + (size_t)(ny - 1 - y) * nx * 3, (size_t)nx * 3);
  return write_jpeg(path, sheet.data(), w, ny, quality);
}
} // namespace tpt

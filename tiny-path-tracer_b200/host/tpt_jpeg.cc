// host/tpt_jpeg.cc -- baseline JPEG decoder for texture input (earthmap.jpg).
//
// The reference decodes its texture with the vendored stb_image v2.23 (third_party/stb_image.h,
// public domain; call sites src/utils.cc:236-240,400). That dependency is not part of this
// repository; this file restates the PUBLISHED algorithm it implements for the case the
// reference's asset needs -- ITU T.81 baseline sequential DCT, 8-bit, Huffman coded, 1 or 3
// components without chroma subsampling -- with the same integer arithmetic (the IJG "islow"
// inverse DCT with 12-bit constants and the 20-bit fixed-point YCbCr conversion), so the decoded
// bytes, which are what crosses the C-ABI as tpt_image_desc, are the ones the reference samples.
// tests/test_host_frontend.py checks that byte for byte against the reference's own decode.
// Anything outside that subset (progressive, subsampled chroma, 12-bit, arithmetic coding) makes
// the decoder return false; load_image_texture then falls back to ImageMagick.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace tpt {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  // canonical Huffman code (T.81 annex C): codes of length L are consecutive
  int mincode[17], maxcode[18], valptr[17];
  uint8_t vals[256];
  bool present = false;
  void build(const uint8_t counts[16], const uint8_t *symbols, int n) {
    std::memcpy(vals, symbols, n);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
      valptr[len] = k;
      mincode[len] = code;
      code += counts[len - 1];
      k += counts[len - 1];
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

struct BitReader {
  const uint8_t *p, *end;
  uint32_t acc = 0;
  int nbits = 0;
  bool hit_marker = false;
  int bit() {
    if (nbits == 0) {
      int b = 0;
      if (!hit_marker && p < end) {
        b = *p++;
        if (b == 0xFF) {
          int b2 = p < end ? *p : 0;
          if (b2 == 0) p++;            // stuffed zero
          else { hit_marker = true; b = 0; p--; } // a marker: feed zeros from here on
        }
      }
      acc = (uint32_t)b;
      nbits = 8;
    }
    nbits--;
    return (acc >> nbits) & 1;
  }
  int bits(int n) {
    int v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | bit();
    return v;
  }
  void reset() { nbits = 0; acc = 0; hit_marker = false; }
};

int decode_symbol(BitReader &br, const Huff &h) {
  int code = 0;
  for (int len = 1; len <= 16; len++) {
    code = (code << 1) | br.bit();
    if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len])
      return h.vals[h.valptr[len] + code - h.mincode[len]];
  }
  return -1;
}

// T.81 F.2.2.1 EXTEND
int extend(int v, int s) { return s == 0 ? 0 : (v < (1 << (s - 1)) ? v - (1 << s) + 1 : v); }

inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// IJG jidctint ("islow") 1-D kernel with 12-bit constants
#define TPT_F2F(x) ((int)(((x) * 4096 + 0.5)))
#define TPT_FSH(x) ((x) * 4096)
#define TPT_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                \
  int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                          \
  p2 = s2;                                                                                         \
  p3 = s6;                                                                                         \
  p1 = (p2 + p3) * TPT_F2F(0.5411961f);                                                            \
  t2 = p1 + p3 * TPT_F2F(-1.847759065f);                                                           \
  t3 = p1 + p2 * TPT_F2F(0.765366865f);                                                            \
  p2 = s0;                                                                                         \
  p3 = s4;                                                                                         \
  t0 = TPT_FSH(p2 + p3);                                                                           \
  t1 = TPT_FSH(p2 - p3);                                                                           \
  x0 = t0 + t3;                                                                                    \
  x3 = t0 - t3;                                                                                    \
  x1 = t1 + t2;                                                                                    \
  x2 = t1 - t2;                                                                                    \
  t0 = s7;                                                                                         \
  t1 = s5;                                                                                         \
  t2 = s3;                                                                                         \
  t3 = s1;                                                                                         \
  p3 = t0 + t2;                                                                                    \
  p4 = t1 + t3;                                                                                    \
  p1 = t0 + t3;                                                                                    \
  p2 = t1 + t2;                                                                                    \
  p5 = (p3 + p4) * TPT_F2F(1.175875602f);                                                          \
  t0 = t0 * TPT_F2F(0.298631336f);                                                                 \
  t1 = t1 * TPT_F2F(2.053119869f);                                                                 \
  t2 = t2 * TPT_F2F(3.072711026f);                                                                 \
  t3 = t3 * TPT_F2F(1.501321110f);                                                                 \
  p1 = p5 + p1 * TPT_F2F(-0.899976223f);                                                           \
  p2 = p5 + p2 * TPT_F2F(-2.562915447f);                                                           \
  p3 = p3 * TPT_F2F(-1.961570560f);                                                                \
  p4 = p4 * TPT_F2F(-0.390180644f);                                                                \
  t3 += p1 + p4;                                                                                   \
  t2 += p2 + p3;                                                                                   \
  t1 += p2 + p4;                                                                                   \
  t0 += p1 + p3;

void idct_block(uint8_t *out, int stride, const short d[64]) {
  int val[64];
  for (int i = 0; i < 8; i++) { // columns, keeping two extra bits of precision
    const short *c = d + i;
    int *v = val + i;
    if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
      int dc = c[0] * 4;
      v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
    } else {
      TPT_IDCT_1D(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56])
      x0 += 512; x1 += 512; x2 += 512; x3 += 512;
      v[0] = (x0 + t3) >> 10;
      v[56] = (x0 - t3) >> 10;
      v[8] = (x1 + t2) >> 10;
      v[48] = (x1 - t2) >> 10;
      v[16] = (x2 + t1) >> 10;
      v[40] = (x2 - t1) >> 10;
      v[24] = (x3 + t0) >> 10;
      v[32] = (x3 - t0) >> 10;
    }
  }
  for (int i = 0; i < 8; i++) { // rows: remove 1<<17, round, re-centre on 128
    const int *v = val + 8 * i;
    uint8_t *o = out + i * stride;
    TPT_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
    x0 += 65536 + (128 << 17);
    x1 += 65536 + (128 << 17);
    x2 += 65536 + (128 << 17);
    x3 += 65536 + (128 << 17);
    o[0] = clamp8((x0 + t3) >> 17);
    o[7] = clamp8((x0 - t3) >> 17);
    o[1] = clamp8((x1 + t2) >> 17);
    o[6] = clamp8((x1 - t2) >> 17);
    o[2] = clamp8((x2 + t1) >> 17);
    o[5] = clamp8((x2 - t1) >> 17);
    o[3] = clamp8((x3 + t0) >> 17);
    o[4] = clamp8((x3 - t0) >> 17);
  }
}

#define TPT_FLOAT2FIXED(x) (((int)((x) * 4096.0f + 0.5f)) << 8)

} // namespace

bool decode_baseline_jpeg(const std::vector<uint8_t> &file, std::vector<uint8_t> &rgb, int &width, int &height) {
  const uint8_t *d = file.data();
  size_t n = file.size(), i = 2;
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
  uint16_t quant[4][64];
  bool have_q[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  int ncomp = 0, comp_id[3], comp_h[3], comp_v[3], comp_tq[3], comp_td[3] = {0, 0, 0}, comp_ta[3] = {0, 0, 0};
  int restart = 0;
  width = height = 0;
  while (i + 4 <= n) {
    if (d[i] != 0xFF) return false;
    int m = d[i + 1];
    if (m == 0xFF) { i++; continue; }
    int L = (d[i + 2] << 8) | d[i + 3];
    const uint8_t *s = d + i + 4;
    if (i + 2 + L > n) return false;
    if (m == 0xDB) { // DQT
      int left = L - 2;
      while (left > 0) {
        int pq = s[0] >> 4, tq = s[0] & 15;
        if (tq > 3) return false;
        s++; left--;
        for (int k = 0; k < 64; k++) {
          quant[tq][kZigzag[k]] = pq ? (uint16_t)((s[0] << 8) | s[1]) : s[0];
          s += pq ? 2 : 1;
          left -= pq ? 2 : 1;
        }
        have_q[tq] = true;
      }
    } else if (m == 0xC4) { // DHT
      int left = L - 2;
      while (left > 0) {
        int tc = s[0] >> 4, th = s[0] & 15;
        if (th > 3 || tc > 1) return false;
        int total = 0;
        for (int k = 0; k < 16; k++) total += s[1 + k];
        if (total > 256) return false;
        (tc ? ac[th] : dc[th]).build(s + 1, s + 17, total);
        s += 17 + total;
        left -= 17 + total;
      }
    } else if (m == 0xC0) { // SOF0: baseline
      if (s[0] != 8) return false;
      height = (s[1] << 8) | s[2];
      width = (s[3] << 8) | s[4];
      ncomp = s[5];
      if ((ncomp != 1 && ncomp != 3) || width <= 0 || height <= 0) return false;
      for (int c = 0; c < ncomp; c++) {
        comp_id[c] = s[6 + 3 * c];
        comp_h[c] = s[7 + 3 * c] >> 4;
        comp_v[c] = s[7 + 3 * c] & 15;
        comp_tq[c] = s[8 + 3 * c];
        if (comp_h[c] != 1 || comp_v[c] != 1 || comp_tq[c] > 3) return false; // no chroma subsampling here
      }
    } else if (m == 0xC1 || m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8)) {
      return false; // extended / progressive / lossless / arithmetic
    } else if (m == 0xDD) {
      restart = (s[0] << 8) | s[1];
    } else if (m == 0xDA) { // SOS: one interleaved scan
      int ns = s[0];
      if (ns != ncomp || ncomp == 0) return false;
      for (int k = 0; k < ns; k++) {
        int id = s[1 + 2 * k], c = -1;
        for (int q = 0; q < ncomp; q++)
          if (comp_id[q] == id) c = q;
        if (c < 0) return false;
        comp_td[c] = s[2 + 2 * k] >> 4;
        comp_ta[c] = s[2 + 2 * k] & 15;
        if (comp_td[c] > 3 || comp_ta[c] > 3 || !dc[comp_td[c]].present || !ac[comp_ta[c]].present || !have_q[comp_tq[c]]) return false;
      }
      const int bw = (width + 7) / 8, bh = (height + 7) / 8;
      std::vector<uint8_t> plane[3];
      for (int c = 0; c < ncomp; c++) plane[c].assign((size_t)bw * 8 * bh * 8, 0);
      BitReader br;
      br.p = d + i + 2 + L;
      br.end = d + n;
      int pred[3] = {0, 0, 0}, todo = restart ? restart : 0x7fffffff;
      for (int by = 0; by < bh; by++) {
        for (int bx = 0; bx < bw; bx++) {
          for (int c = 0; c < ncomp; c++) {
            short blk[64];
            std::memset(blk, 0, sizeof(blk));
            int t = decode_symbol(br, dc[comp_td[c]]);
            if (t < 0 || t > 15) return false;
            int diff = t ? extend(br.bits(t), t) : 0;
            pred[c] += diff;
            blk[0] = (short)(pred[c] * quant[comp_tq[c]][0]);
            for (int k = 1; k < 64;) {
              int rs = decode_symbol(br, ac[comp_ta[c]]);
              if (rs < 0) return false;
              int r = rs >> 4, sz = rs & 15;
              if (sz == 0) {
                if (rs != 0xF0) break; // EOB
                k += 16;
              } else {
                k += r;
                if (k > 63) return false;
                int z = kZigzag[k];
                blk[z] = (short)(extend(br.bits(sz), sz) * quant[comp_tq[c]][z]);
                k++;
              }
            }
            idct_block(plane[c].data() + (size_t)by * 8 * bw * 8 + bx * 8, bw * 8, blk);
          }
          if (--todo <= 0) { // restart interval: byte align, skip RSTn, reset predictors
            br.reset();
            while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) br.p++;
            if (br.p + 1 < br.end) br.p += 2;
            pred[0] = pred[1] = pred[2] = 0;
            todo = restart;
          }
        }
      }
      rgb.resize((size_t)width * height * 3);
      const int stride = bw * 8;
      for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
          uint8_t *o = &rgb[((size_t)y * width + x) * 3];
          if (ncomp == 1) {
            o[0] = o[1] = o[2] = plane[0][(size_t)y * stride + x];
          } else { // 20-bit fixed-point YCbCr -> RGB
            int yf = (plane[0][(size_t)y * stride + x] << 20) + (1 << 19);
            int cb = plane[1][(size_t)y * stride + x] - 128, cr = plane[2][(size_t)y * stride + x] - 128;
            int r = yf + cr * TPT_FLOAT2FIXED(1.40200f);
            int g = yf + (cr * -TPT_FLOAT2FIXED(0.71414f)) + ((cb * -TPT_FLOAT2FIXED(0.34414f)) & 0xffff0000);
            int b = yf + cb * TPT_FLOAT2FIXED(1.77200f);
            o[0] = clamp8(r >> 20);
            o[1] = clamp8(g >> 20);
            o[2] = clamp8(b >> 20);
          }
        }
      }
      return true;
    } else if (m == 0xD9) {
      return false;
    }
    i += 2 + L;
  }
  return false;
}

bool read_jpeg(const std::string &path, std::vector<uint8_t> &rgb, int &w, int &h) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<uint8_t> bytes;
  uint8_t buf[65536];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + got);
  std::fclose(f);
  return decode_baseline_jpeg(bytes, rgb, w, h);
}

} // namespace tpt

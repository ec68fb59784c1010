// jpeg_baseline.cpp — decoder for the texture files the reference's scene functions open at run time
// (textures/<name>.jpg through image_io.h:24-41 -> stbi_load(path, &w, &h, &n, 3); main.cu:816, 1010, 1186, 1248, 1254).
//
// The reference decodes with the stb_image single-header library vendored under external/ (a third-party dependency,
// not reference code; v2.x). An image texture's texels enter the render bit for bit (texture.cuh:66-92), so "the same
// image" means the same BYTES, and two conforming JPEG decoders differ in exactly the places the standard leaves open:
// the inverse DCT's integer arithmetic, chroma upsampling and the YCbCr -> RGB fixed-point conversion. This file is an
// independent baseline (SOF0, Huffman, 8-bit, 1 or 3 components, h/v sampling factors 1 or 2, restart intervals)
// decoder that makes the same choices as that library, restated from its published algorithm:
//   * dequantised coefficients as 16-bit, inverse DCT = the 12-bit fixed-point LL&M variant with a column pass rounded
//     at >> 10 and a row pass at >> 17 (+128 level shift), results clamped to [0, 255];
//   * chroma upsampling h2v2 = the "3/4 near + 1/4 far" triangle filter applied vertically then horizontally
//     (9-3-3-1), rounding +8 >> 4 inside a row, +2 >> 2 at the row ends; h2v1 / h1v2 = the 1-D version;
//   * YCbCr -> RGB in 20-bit fixed point with the constants rounded to 12 bits first.
// tests/test_host.py checks the output against the PPMs that the reference's own decoder produced from the same files
// (made once where the reference sources exist, by the checker-side harness): byte-identical for all five textures.
// Progressive (SOF2), arithmetic-coded, 12-bit and CMYK files are rejected with an error: the reference's texture set
// has none, and a silently different decode would break parity.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "scene_builder.h"

namespace rt {
namespace {

const uint8_t kDezigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huff {
  // canonical code tables (ITU T.81 Annex C / F.2.2.3)
  uint8_t size[257];
  uint16_t code[256];
  uint8_t values[256];
  int maxcode[18];   // first code of each length, left-aligned to 16 bits, exclusive upper bound
  int delta[17];     // value index = code - delta[len]
  bool build(const int* count) {
    int k = 0;
    for (int i = 0; i < 16; ++i)
      for (int j = 0; j < count[i]; ++j) { if (k >= 256) return false; size[k++] = (uint8_t)(i + 1); }
    size[k] = 0;
    int c = 0;
    k = 0;
    for (int j = 1; j <= 16; ++j) {
      delta[j] = k - c;
      if (size[k] == j) {
        while (size[k] == j) code[k++] = (uint16_t)(c++);
        if (c - 1 >= (1 << j)) return false;
      }
      maxcode[j] = c << (16 - j);
      c <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    return true;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0;
  int dc_pred = 0;
  int x = 0, y = 0;     // size in samples
  int w2 = 0, h2 = 0;   // size of the stored plane (whole MCUs)
  std::vector<uint8_t> data;
};

struct Decoder {
  const uint8_t* p; const uint8_t* end;
  std::string err;
  Huff hdc[4], hac[4];
  uint16_t dequant[4][64];
  Component comp[3];
  int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1, restart_interval = 0;
  // entropy-coded segment reader
  uint32_t code_buffer = 0; int code_bits = 0; int marker = 0xff; bool nomore = false; int todo = 0;

  bool fail(const char* m) { if (err.empty()) err = m; return false; }
  int get8() { return p < end ? *p++ : 0; }
  int get16() { const int a = get8(); return (a << 8) | get8(); }

  void grow() {
    do {
      unsigned b = nomore ? 0 : (unsigned)get8();
      if (b == 0xff) {
        int c = get8();
        while (c == 0xff) c = get8();  // fill bytes
        if (c != 0) { marker = c; nomore = true; return; }
      }
      code_buffer |= b << (24 - code_bits);
      code_bits += 8;
    } while (code_bits <= 24);
  }
  int decode(const Huff& h) {
    if (code_bits < 16) grow();
    const int temp = (int)(code_buffer >> 16);
    int k = 1;
    while (k <= 16 && temp >= h.maxcode[k]) ++k;
    if (k == 17 || k > code_bits) { code_bits -= 16; return -1; }
    const int c = (int)((code_buffer >> (32 - k)) & ((1u << k) - 1)) + h.delta[k];
    if (c < 0 || c >= 256) return -1;
    code_bits -= k;
    code_buffer <<= k;
    return h.values[c];
  }
  // receive n bits and sign-extend (T.81 F.2.2.1 EXTEND)
  int extend_receive(int n) {
    if (n == 0) return 0;
    if (code_bits < n) grow();
    if (code_bits < n) return 0;
    const int v = (int)(code_buffer >> (32 - n));
    code_buffer <<= n;
    code_bits -= n;
    return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
  }
  bool decode_block(short data[64], Component& c) {
    const int t = decode(hdc[c.hd]);
    if (t < 0 || t > 15) return fail("bad huffman code");
    memset(data, 0, 64 * sizeof(short));
    const int diff = t ? extend_receive(t) : 0;
    const int dc = c.dc_pred + diff;
    c.dc_pred = dc;
    const uint16_t* dq = dequant[c.tq];
    data[0] = (short)(dc * dq[0]);
    int k = 1;
    do {
      const int rs = decode(hac[c.ha]);
      if (rs < 0) return fail("bad huffman code");
      const int s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (rs != 0xf0) break;  // end of block
        k += 16;
      } else {
        k += r;
        const int zig = kDezigzag[k++];
        data[zig] = (short)(extend_receive(s) * dq[zig]);
      }
    } while (k < 64);
    return true;
  }

  static uint8_t clamp8(int x) { return (unsigned)x > 255u ? (x < 0 ? 0 : 255) : (uint8_t)x; }

#define RT_F2F(x) ((int)(((x) * 4096 + 0.5)))
#define RT_FSH(x) ((x) * 4096)
#define RT_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                                     \
  int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                                              \
  p2 = s2; p3 = s6;                                                                                                    \
  p1 = (p2 + p3) * RT_F2F(0.5411961f);                                                                                 \
  t2 = p1 + p3 * RT_F2F(-1.847759065f);                                                                                \
  t3 = p1 + p2 * RT_F2F(0.765366865f);                                                                                 \
  p2 = s0; p3 = s4;                                                                                                    \
  t0 = RT_FSH(p2 + p3); t1 = RT_FSH(p2 - p3);                                                                          \
  x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                                              \
  t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                                                  \
  p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                                                              \
  p5 = (p3 + p4) * RT_F2F(1.175875602f);                                                                               \
  t0 = t0 * RT_F2F(0.298631336f); t1 = t1 * RT_F2F(2.053119869f);                                                      \
  t2 = t2 * RT_F2F(3.072711026f); t3 = t3 * RT_F2F(1.501321110f);                                                      \
  p1 = p5 + p1 * RT_F2F(-0.899976223f); p2 = p5 + p2 * RT_F2F(-2.562915447f);                                          \
  p3 = p3 * RT_F2F(-1.961570560f); p4 = p4 * RT_F2F(-0.390180644f);                                                    \
  t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

  static void idct_block(uint8_t* out, int out_stride, const short data[64]) {
    int val[64];
    int* v = val;
    const short* d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {
      if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
        const int dcterm = d[0] * 4;
        v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
      } else {
        RT_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
        x0 += 512; x1 += 512; x2 += 512; x3 += 512;
        v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
        v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
        v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
        v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
      }
    }
    v = val;
    uint8_t* o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += out_stride) {
      RT_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
      x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
      o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
      o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
      o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
      o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
  }
#undef RT_IDCT_1D
#undef RT_F2F
#undef RT_FSH

  void reset_entropy() {
    code_bits = 0; code_buffer = 0; nomore = false; marker = 0xff;
    for (int i = 0; i < ncomp; ++i) comp[i].dc_pred = 0;
    todo = restart_interval ? restart_interval : 0x7fffffff;
  }

  bool parse_scan_data(const int* order, int ns) {
    reset_entropy();
    short blk[64];
    if (ns == 1) {
      Component& c = comp[order[0]];
      const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
      for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
          if (!decode_block(blk, c)) return false;
          idct_block(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, blk);
          if (--todo <= 0) { if (!restart()) return true; }
        }
      return true;
    }
    const int mcu_w = hmax * 8, mcu_h = vmax * 8;
    const int mx = (width + mcu_w - 1) / mcu_w, my = (height + mcu_h - 1) / mcu_h;
    for (int j = 0; j < my; ++j)
      for (int i = 0; i < mx; ++i) {
        for (int k = 0; k < ns; ++k) {
          Component& c = comp[order[k]];
          for (int y = 0; y < c.v; ++y)
            for (int x = 0; x < c.h; ++x) {
              const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
              if (!decode_block(blk, c)) return false;
              idct_block(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, blk);
            }
        }
        if (--todo <= 0) { if (!restart()) return true; }
      }
    return true;
  }
  // at a restart interval boundary: expect RSTn, reset the predictors; false = no restart marker (end of data)
  bool restart() {
    if (code_bits < 24) grow();
    if (!(marker >= 0xd0 && marker <= 0xd7)) return false;
    reset_entropy();
    return true;
  }

  bool parse_headers_and_scans() {
    if (get8() != 0xff || get8() != 0xd8) return fail("not a JPEG file");
    bool have_frame = false;
    int pending = -1;
    for (;;) {
      int m;
      if (pending >= 0) { m = pending; pending = -1; }
      else {
        m = get8();
        while (m != 0xff) { if (p >= end) return fail("no end-of-image marker"); m = get8(); }
        while (m == 0xff) m = get8();
      }
      if (m == 0xd9) break;  // EOI
      if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
      int len = get16();
      if (len < 2 || p + (len - 2) > end) return fail("corrupt marker segment");
      const uint8_t* seg_end = p + (len - 2);
      if (m == 0xdb) {  // DQT
        while (p < seg_end) {
          const int q = get8(), prec = q >> 4, t = q & 15;
          if (prec > 1 || t > 3) return fail("bad DQT");
          for (int i = 0; i < 64; ++i) dequant[t][kDezigzag[i]] = (uint16_t)(prec ? get16() : get8());
        }
      } else if (m == 0xc4) {  // DHT
        while (p < seg_end) {
          const int q = get8(), tc = q >> 4, th = q & 15;
          if (tc > 1 || th > 3) return fail("bad DHT");
          int count[16], n = 0;
          for (int i = 0; i < 16; ++i) { count[i] = get8(); n += count[i]; }
          if (n > 256) return fail("bad DHT");
          Huff& h = tc ? hac[th] : hdc[th];
          if (!h.build(count)) return fail("bad huffman code lengths");
          for (int i = 0; i < n; ++i) h.values[i] = (uint8_t)get8();
        }
      } else if (m == 0xdd) {  // DRI
        restart_interval = get16();
      } else if (m == 0xc0 || m == 0xc1) {  // SOF0 / SOF1 (Huffman, sequential)
        if (get8() != 8) return fail("only 8-bit JPEG is supported");
        height = get16(); width = get16();
        ncomp = get8();
        if (width <= 0 || height <= 0) return fail("bad image size");
        if (ncomp != 1 && ncomp != 3) return fail("only grey and YCbCr JPEG are supported");
        for (int i = 0; i < ncomp; ++i) {
          comp[i].id = get8();
          const int q = get8();
          comp[i].h = q >> 4; comp[i].v = q & 15; comp[i].tq = get8();
          if (comp[i].h < 1 || comp[i].h > 2 || comp[i].v < 1 || comp[i].v > 2 || comp[i].tq > 3) return fail("unsupported sampling factors");
          if (comp[i].h > hmax) hmax = comp[i].h;
          if (comp[i].v > vmax) vmax = comp[i].v;
        }
        const int mcu_w = hmax * 8, mcu_h = vmax * 8;
        const int mx = (width + mcu_w - 1) / mcu_w, my = (height + mcu_h - 1) / mcu_h;
        for (int i = 0; i < ncomp; ++i) {
          Component& c = comp[i];
          c.x = (width * c.h + hmax - 1) / hmax; c.y = (height * c.v + vmax - 1) / vmax;
          c.w2 = mx * c.h * 8; c.h2 = my * c.v * 8;
          c.data.assign((size_t)c.w2 * c.h2, 0);
        }
        have_frame = true;
      } else if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
        return fail("progressive / lossless / arithmetic-coded JPEG is not supported (baseline files only)");
      } else if (m == 0xda) {  // SOS
        if (!have_frame) return fail("scan before frame header");
        const int ns = get8();
        if (ns < 1 || ns > ncomp) return fail("bad SOS");
        int order[3];
        for (int i = 0; i < ns; ++i) {
          const int id = get8(), q = get8();
          int which = -1;
          for (int k = 0; k < ncomp; ++k) if (comp[k].id == id) which = k;
          if (which < 0) return fail("bad SOS component");
          comp[which].hd = q >> 4; comp[which].ha = q & 15;
          if (comp[which].hd > 3 || comp[which].ha > 3) return fail("bad SOS table");
          order[i] = which;
        }
        p = seg_end;
        if (!parse_scan_data(order, ns)) return false;
        if (nomore && marker != 0xff) pending = marker;  // the bit reader ran into the next marker
        continue;
      }
      p = seg_end;
    }
    return have_frame;
  }
};

// chroma upsampling rows (near = the sample row closer to the output row)
void row_1(uint8_t* out, const uint8_t* n, const uint8_t*, int w) { memcpy(out, n, (size_t)w); }
void row_v2(uint8_t* out, const uint8_t* n, const uint8_t* f, int w) { for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * n[i] + f[i] + 2) >> 2); }
void row_h2(uint8_t* out, const uint8_t* in, const uint8_t*, int w) {
  if (w == 1) { out[0] = out[1] = in[0]; return; }
  out[0] = in[0];
  out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
  int i;
  for (i = 1; i < w - 1; ++i) {
    const int n = 3 * in[i] + 2;
    out[i * 2 + 0] = (uint8_t)((n + in[i - 1]) >> 2);
    out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
  }
  out[i * 2 + 0] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
  out[i * 2 + 1] = in[w - 1];
}
void row_hv2(uint8_t* out, const uint8_t* n, const uint8_t* f, int w) {
  if (w == 1) { out[0] = out[1] = (uint8_t)((3 * n[0] + f[0] + 2) >> 2); return; }
  int t1 = 3 * n[0] + f[0];
  out[0] = (uint8_t)((t1 + 2) >> 2);
  for (int i = 1; i < w; ++i) {
    const int t0 = t1;
    t1 = 3 * n[i] + f[i];
    out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
    out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
  }
  out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
}

}  // namespace

// Decodes `path` to 8-bit RGB (grey files are replicated into three channels, like stbi_load(..., 3)).
bool load_jpeg(const std::string& path, HostImage& out, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  Decoder d;
  d.p = buf.data(); d.end = buf.data() + buf.size();
  memset(d.dequant, 0, sizeof(d.dequant));
  if (!d.parse_headers_and_scans()) { err = path + ": " + (d.err.empty() ? "corrupt JPEG" : d.err); return false; }
  const int W = d.width, H = d.height;
  out.width = W; out.height = H; out.bpp = 3;
  out.px.assign((size_t)W * H * 3, 0);
  struct Res { int hs, vs, ystep, w_lores, ypos; const uint8_t *line0, *line1; std::vector<uint8_t> linebuf; } res[3];
  for (int k = 0; k < d.ncomp; ++k) {
    Res& r = res[k];
    r.hs = d.hmax / d.comp[k].h; r.vs = d.vmax / d.comp[k].v;
    r.ystep = r.vs >> 1;
    r.w_lores = (W + r.hs - 1) / r.hs;
    r.ypos = 0;
    r.line0 = r.line1 = d.comp[k].data.data();
    r.linebuf.assign((size_t)W + 3, 0);
  }
  const uint8_t* co[3] = {nullptr, nullptr, nullptr};
  for (int j = 0; j < H; ++j) {
    uint8_t* o = out.px.data() + (size_t)3 * W * j;
    for (int k = 0; k < d.ncomp; ++k) {
      Res& r = res[k];
      const bool y_bot = r.ystep >= (r.vs >> 1);
      const uint8_t* nr = y_bot ? r.line1 : r.line0;
      const uint8_t* fr = y_bot ? r.line0 : r.line1;
      if (r.hs == 1 && r.vs == 1) co[k] = nr;
      else {
        if (r.hs == 1 && r.vs == 2) row_v2(r.linebuf.data(), nr, fr, r.w_lores);
        else if (r.hs == 2 && r.vs == 1) row_h2(r.linebuf.data(), nr, fr, r.w_lores);
        else row_hv2(r.linebuf.data(), nr, fr, r.w_lores);
        co[k] = r.linebuf.data();
      }
      if (++r.ystep >= r.vs) {
        r.ystep = 0;
        r.line0 = r.line1;
        if (++r.ypos < d.comp[k].y) r.line1 += d.comp[k].w2;
      }
    }
    if (d.ncomp == 1) {
      for (int i = 0; i < W; ++i) o[3 * i] = o[3 * i + 1] = o[3 * i + 2] = co[0][i];
    } else {
      // YCbCr -> RGB: 20-bit fixed point, constants rounded to 12 bits first
#define RT_FX(x) (((int)((x) * 4096.0f + 0.5f)) << 8)
      for (int i = 0; i < W; ++i) {
        const int y_fixed = (co[0][i] << 20) + (1 << 19);
        const int cr = co[2][i] - 128, cb = co[1][i] - 128;
        int r = y_fixed + cr * RT_FX(1.40200f);
        int g = y_fixed + (cr * -RT_FX(0.71414f)) + (int)(((unsigned)(cb * -RT_FX(0.34414f))) & 0xffff0000u);
        int b = y_fixed + cb * RT_FX(1.77200f);
        r >>= 20; g >>= 20; b >>= 20;
        o[3 * i] = Decoder::clamp8(r); o[3 * i + 1] = Decoder::clamp8(g); o[3 * i + 2] = Decoder::clamp8(b);
      }
#undef RT_FX
    }
  }
  return true;
}

}  // namespace rt

// File decode in front of image_ops::preprocess_image / load_image_as_tensor: `image::open(file)?.into_rgba()`
// (image_ops.rs:193) and `.into_luma()` (:78) — SURVEY §8 f4.  Compiled with -fmad=false (the colour conversion
// must round like the reference's scalar f32 code).
//
// Split of the work (B200-first): the entropy-coded part of a file is a serial bit stream, so it is decoded on the
// host, one worker thread per image (JPEG: Huffman -> DCT coefficients, ITU-T T.81 Annex F / G, baseline and
// progressive; PNG: inflate + the five row filters); everything that is data-parallel runs on the device for the
// whole batch in two launches:
//   jpeg_idct_kernel      dequantisation + 8x8 inverse DCT (jpeg-decoder 0.1.20 idct.rs: stb-style integer
//                         butterflies, 12-bit constants), 8 threads per block, 4 blocks per warp, coefficients read
//                         once (128 B per block), component planes written once with 8-byte stores
//   decode_assemble_kernel  chroma upsampling (upsampler.rs H2V1 / H1V2 / H2V2 triangle filters in closed form,
//                         on the component's real size) + YCbCr -> RGB (decoder.rs, f32) -> RGBA8 or luma;
//                         PNG pixels go through the same kernel (L / LA / RGB / RGBA -> RGBA8 or luma)
// The RGBA arena it writes is exactly what ocrb_preprocess_rgba_batch reads, so ocrb_preprocess_files is
// preprocess_image(file, dims) for a batch with the decoded pixels never leaving HBM.
#include <atomic>
#include <memory>
#include <thread>

#include "common.cuh"

namespace ocrb {

// ===================================================================================================================
// host: JPEG entropy decoding
// ===================================================================================================================
namespace {

const uint8_t kNatural[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                              15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// MSB-first bit reader over an entropy-coded segment: removes FF00 stuffing, stops at a marker (zeros after it)
struct JBits {
  const uint8_t *p, *end;
  uint64_t acc = 0;
  int n = 0, marker = 0;
  JBits(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
  void refill() {
    while (n <= 56) {
      uint64_t b = 0;
      if (!marker && p < end) {
        b = *p++;
        if (b == 0xFF) {
          while (p < end && *p == 0xFF) ++p;
          const int m = p < end ? *p++ : 0xD9;
          if (m != 0) { marker = m; b = 0; }
        }
      }
      acc |= b << (56 - n);
      n += 8;
    }
  }
  inline int peek(int k) { if (n < k) refill(); return (int)(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; n -= k; }
  inline int get(int k) { if (k == 0) return 0; const int v = peek(k); skip(k); return v; }
  inline int bit() { return get(1); }
  // restart boundary: drop the padding bits and step over the RSTn marker
  bool restart() {
    acc = 0;
    n = 0;
    if (!marker) {
      while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
      if (p + 1 >= end) return false;
      p += 2;
      return true;
    }
    if (marker < 0xD0 || marker > 0xD7) return false;
    marker = 0;
    return true;
  }
};

// canonical Huffman table: 9-bit direct lookup, longer codes by the per-length bounds
struct JHuff {
  static constexpr int FAST = 9;
  uint16_t fast[1 << FAST];  // (length << 8) | symbol, 0 = not a short code
  int32_t maxcode[18];       // left-aligned to 16 bits, exclusive upper bound per length
  int32_t delta[17];
  uint8_t sym[256];
  bool present = false;
  bool build(const uint8_t *counts, const uint8_t *symbols, int total) {
    memset(fast, 0, sizeof fast);
    memcpy(sym, symbols, (size_t)total);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
      delta[len] = k - code;
      for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
        if (len <= FAST) {
          const int lo = code << (FAST - len);
          for (int f = 0; f < (1 << (FAST - len)); ++f) fast[lo + f] = (uint16_t)((len << 8) | symbols[k]);
        }
      }
      if (code > (1 << len)) return false;
      maxcode[len] = code << (16 - len);
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
    return true;
  }
  inline int decode(JBits &b) const {
    const int look = b.peek(16);
    const uint16_t f = fast[look >> (16 - FAST)];
    if (f) { b.skip(f >> 8); return f & 255; }
    int len = FAST + 1;
    while (look >= maxcode[len]) ++len;
    if (len > 16) return -1;
    b.skip(len);
    return sym[((look >> (16 - len)) + delta[len]) & 255];
  }
};

inline int jextend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

struct JComp {
  int id = 0, h = 1, v = 1, tq = 0;
  int w = 0, hpx = 0;    // real plane size: ceil(image size * factor / max factor)
  int bw = 0, bh = 0;    // block grid of the plane buffer (whole MCUs)
  int64_t coef_off = 0;  // int16 elements into the batch coefficient arena
  int64_t plane_off = 0; // bytes into the batch plane arena
  int pred = 0, td = 0, ta = 0;
};

struct JFrame {
  int W = 0, H = 0, nc = 0, hmax = 1, vmax = 1;
  bool progressive = false;
  JComp comp[3];
  uint16_t qt[4][64];
  bool qt_ok[4] = {false, false, false, false};
};

// reads the markers up to and including the frame header: size, components, sampling
int jpeg_parse_frame(const uint8_t *d, size_t n, JFrame *f) {
  size_t i = 2;
  while (i + 4 <= n) {
    if (d[i] != 0xFF) return OCRB_ERR_INVALID;
    while (i < n && d[i] == 0xFF) ++i;
    if (i >= n) break;
    const int m = d[i++];
    if (m == 0xD9 || m == 0xDA) break;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (i + 2 > n) break;
    const size_t L = ((size_t)d[i] << 8) | d[i + 1];
    if (L < 2 || i + L > n) break;
    const uint8_t *s = d + i + 2;
    if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
      if (L < 8 || s[0] != 8) { set_error("jpeg: only 8-bit samples are supported"); return OCRB_ERR_INVALID; }
      f->H = (s[1] << 8) | s[2];
      f->W = (s[3] << 8) | s[4];
      f->nc = s[5];
      f->progressive = m == 0xC2;
      if (f->W == 0 || f->H == 0 || (f->nc != 1 && f->nc != 3) || L < 8 + 3 * (size_t)f->nc) {
        set_error("jpeg: unsupported frame (%d components, %dx%d)", f->nc, f->W, f->H);
        return OCRB_ERR_INVALID;
      }
      for (int c = 0; c < f->nc; ++c) {
        JComp &k = f->comp[c];
        k.id = s[6 + 3 * c];
        k.h = s[7 + 3 * c] >> 4;
        k.v = s[7 + 3 * c] & 15;
        k.tq = s[8 + 3 * c];
        if (k.h < 1 || k.h > 4 || k.v < 1 || k.v > 4 || k.tq > 3) { set_error("jpeg: bad component header"); return OCRB_ERR_INVALID; }
        if (f->nc == 1) k.h = k.v = 1;
        f->hmax = k.h > f->hmax ? k.h : f->hmax;
        f->vmax = k.v > f->vmax ? k.v : f->vmax;
      }
      const int mx = (f->W + 8 * f->hmax - 1) / (8 * f->hmax), my = (f->H + 8 * f->vmax - 1) / (8 * f->vmax);
      for (int c = 0; c < f->nc; ++c) {
        JComp &k = f->comp[c];
        if ((k.h != f->hmax && 2 * k.h != f->hmax) || (k.v != f->vmax && 2 * k.v != f->vmax)) {
          set_error("jpeg: sampling ratio %dx%d of %dx%d is not supported (the reference's decoder handles 1 and 2)", k.h, k.v, f->hmax, f->vmax);
          return OCRB_ERR_INVALID;
        }
        k.w = (f->W * k.h + f->hmax - 1) / f->hmax;
        k.hpx = (f->H * k.v + f->vmax - 1) / f->vmax;
        k.bw = mx * k.h;
        k.bh = my * k.v;
      }
      return OCRB_OK;
    }
    if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC && m != 0xC4)) {
      set_error("jpeg: coding process SOF%d is not supported", m - 0xC0);
      return OCRB_ERR_INVALID;
    }
    i += L;
  }
  set_error("jpeg: no frame header");
  return OCRB_ERR_INVALID;
}

struct JScan {
  int ss = 0, se = 63, ah = 0, al = 0, eobrun = 0;
};

inline bool block_sequential(JBits &b, JComp &c, const JHuff &dc, const JHuff &ac, int16_t *q) {
  int s = dc.decode(b);
  if (s < 0 || s > 11) return false;
  c.pred += jextend(b.get(s), s);
  q[0] = (int16_t)c.pred;
  int k = 1;
  while (k < 64) {
    const int rs = ac.decode(b);
    if (rs < 0) return false;
    s = rs & 15;
    if (s == 0) {
      if (rs != 0xF0) break;
      k += 16;
      continue;
    }
    k += rs >> 4;
    if (k > 63) return false;
    q[kNatural[k++]] = (int16_t)jextend(b.get(s), s);
  }
  return true;
}

inline void refine(JBits &b, int16_t &c, int bitval) {
  if (b.bit() && (c & bitval) == 0) c = (int16_t)(c >= 0 ? c + bitval : c - bitval);
}

bool block_progressive(JBits &b, JComp &c, const JHuff &dc, const JHuff &ac, int16_t *q, JScan &sc) {
  if (sc.ss == 0) {
    if (sc.ah == 0) {
      const int s = dc.decode(b);
      if (s < 0 || s > 11) return false;
      c.pred += jextend(b.get(s), s);
      q[0] = (int16_t)(c.pred * (1 << sc.al));
    } else if (b.bit()) {
      q[0] |= (int16_t)(1 << sc.al);
    }
    return true;
  }
  if (sc.ah == 0) {
    if (sc.eobrun > 0) { --sc.eobrun; return true; }
    int k = sc.ss;
    while (k <= sc.se) {
      const int rs = ac.decode(b);
      if (rs < 0) return false;
      const int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r < 15) {
          sc.eobrun = (1 << r) - 1 + (r ? b.get(r) : 0);
          break;
        }
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) return false;
      q[kNatural[k++]] = (int16_t)(jextend(b.get(s), s) * (1 << sc.al));
    }
    return true;
  }
  // successive-approximation refinement of an AC band (T.81 G.1.2.3)
  const int bitval = 1 << sc.al;
  int k = sc.ss;
  if (sc.eobrun == 0) {
    while (k <= sc.se) {
      const int rs = ac.decode(b);
      if (rs < 0) return false;
      int r = rs >> 4;
      const int s = rs & 15;
      int fresh = 0;
      if (s == 0) {
        if (r < 15) {
          sc.eobrun = (1 << r) + (r ? b.get(r) : 0);
          break;
        }
      } else {
        if (s != 1) return false;
        fresh = b.bit() ? bitval : -bitval;
      }
      while (k <= sc.se) {
        int16_t &v = q[kNatural[k++]];
        if (v != 0) {
          refine(b, v, bitval);
        } else if (r-- == 0) {
          if (fresh) v = (int16_t)fresh;
          break;
        }
      }
    }
  }
  if (sc.eobrun > 0) {
    for (; k <= sc.se; ++k) {
      int16_t &v = q[kNatural[k]];
      if (v != 0) refine(b, v, bitval);
    }
    --sc.eobrun;
  }
  return true;
}

// entropy-decodes every scan of the file into `coef` (zero-initialised by the caller); fills the quantisation tables
int jpeg_decode_coefficients(const uint8_t *d, size_t n, JFrame *f, int16_t *coef, std::string *err) {
  std::unique_ptr<JHuff[]> hdc(new JHuff[4]), hac(new JHuff[4]);
  int restart = 0;
  size_t i = 2;
  auto fail = [&](const char *what) { *err = what; return OCRB_ERR_INVALID; };
  for (;;) {
    if (i + 2 > n) return fail("jpeg: truncated file");
    if (d[i] != 0xFF) return fail("jpeg: marker expected");
    while (i < n && d[i] == 0xFF) ++i;
    if (i >= n) return fail("jpeg: truncated file");
    const int m = d[i++];
    if (m == 0xD9) break;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (i + 2 > n) return fail("jpeg: truncated segment");
    const size_t L = ((size_t)d[i] << 8) | d[i + 1];
    if (L < 2 || i + L > n) return fail("jpeg: truncated segment");
    const uint8_t *s = d + i + 2;
    const size_t sl = L - 2;
    if (m == 0xDB) {
      for (size_t o = 0; o < sl;) {
        const int wide = s[o] >> 4, t = s[o] & 15;
        if (t > 3 || wide > 1 || o + 1 + 64 * (size_t)(wide + 1) > sl) return fail("jpeg: bad quantisation table");
        for (int k = 0; k < 64; ++k) f->qt[t][kNatural[k]] = wide ? (uint16_t)((s[o + 1 + 2 * k] << 8) | s[o + 2 + 2 * k]) : s[o + 1 + k];
        f->qt_ok[t] = true;
        o += 1 + 64 * (size_t)(wide + 1);
      }
    } else if (m == 0xC4) {
      for (size_t o = 0; o < sl;) {
        if (o + 17 > sl) return fail("jpeg: bad Huffman table");
        const int cls = s[o] >> 4, t = s[o] & 15;
        int total = 0;
        for (int l = 0; l < 16; ++l) total += s[o + 1 + l];
        if (cls > 1 || t > 3 || total > 256 || o + 17 + (size_t)total > sl) return fail("jpeg: bad Huffman table");
        if (!(cls ? hac[t] : hdc[t]).build(s + o + 1, s + o + 17, total)) return fail("jpeg: bad Huffman table");
        o += 17 + (size_t)total;
      }
    } else if (m == 0xDD) {
      if (sl < 2) return fail("jpeg: bad restart interval");
      restart = (s[0] << 8) | s[1];
    } else if (m == 0xEE) {
      if (sl >= 12 && memcmp(s, "Adobe", 5) == 0 && s[11] == 0 && f->nc == 3) return fail("jpeg: RGB-coded files are not supported");
    } else if (m == 0xDA) {
      const int ns = sl ? s[0] : 0;
      if (ns < 1 || ns > f->nc || sl < 4 + 2 * (size_t)ns) return fail("jpeg: bad scan header");
      JComp *sc_comp[3];
      for (int k = 0; k < ns; ++k) {
        sc_comp[k] = nullptr;
        for (int c = 0; c < f->nc; ++c)
          if (f->comp[c].id == s[1 + 2 * k]) sc_comp[k] = &f->comp[c];
        if (!sc_comp[k]) return fail("jpeg: scan names an unknown component");
        sc_comp[k]->td = s[2 + 2 * k] >> 4;
        sc_comp[k]->ta = s[2 + 2 * k] & 15;
        if (sc_comp[k]->td > 3 || sc_comp[k]->ta > 3) return fail("jpeg: bad table selector");
      }
      JScan sc;
      if (f->progressive) {
        sc.ss = s[1 + 2 * ns];
        sc.se = s[2 + 2 * ns];
        sc.ah = s[3 + 2 * ns] >> 4;
        sc.al = s[3 + 2 * ns] & 15;
        if (sc.ss > sc.se || sc.se > 63 || (sc.ss > 0 && ns != 1) || (sc.ss == 0 && sc.se != 0) || sc.al > 13) return fail("jpeg: bad progressive scan");
      }
      const bool need_dc = sc.ss == 0 && sc.ah == 0, need_ac = sc.se > 0;
      for (int k = 0; k < ns; ++k)
        if ((need_dc && !hdc[sc_comp[k]->td].present) || (need_ac && !hac[sc_comp[k]->ta].present)) return fail("jpeg: missing Huffman table");
      for (int c = 0; c < f->nc; ++c) f->comp[c].pred = 0;
      JBits b(d + i + L, d + n);
      // a one-component scan walks that component's own block grid; otherwise MCU by MCU
      const int ux = ns == 1 ? (sc_comp[0]->w + 7) / 8 : (f->W + 8 * f->hmax - 1) / (8 * f->hmax);
      const int uy = ns == 1 ? (sc_comp[0]->hpx + 7) / 8 : (f->H + 8 * f->vmax - 1) / (8 * f->vmax);
      int left = restart;
      for (int y = 0; y < uy; ++y)
        for (int x = 0; x < ux; ++x) {
          if (restart && left == 0) {
            if (!b.restart()) return fail("jpeg: restart marker expected");
            for (int c = 0; c < f->nc; ++c) f->comp[c].pred = 0;
            sc.eobrun = 0;
            left = restart;
          }
          for (int k = 0; k < ns; ++k) {
            JComp &c = *sc_comp[k];
            const int nh = ns == 1 ? 1 : c.h, nv = ns == 1 ? 1 : c.v;
            for (int by = 0; by < nv; ++by)
              for (int bx = 0; bx < nh; ++bx) {
                int16_t *q = coef + c.coef_off + ((int64_t)(y * nv + by) * c.bw + (x * nh + bx)) * 64;
                const bool ok = f->progressive ? block_progressive(b, c, hdc[c.td], hac[c.ta], q, sc) : block_sequential(b, c, hdc[c.td], hac[c.ta], q);
                if (!ok) return fail("jpeg: corrupt entropy-coded data");
              }
          }
          if (restart) --left;
        }
      // resume the marker walk behind the entropy-coded segment
      if (b.marker) {
        i = (size_t)(b.p - d) - 2;
        while (i > 0 && d[i] != 0xFF) --i;  // (fill bytes in front of the marker)
      } else {
        const uint8_t *q = b.p;
        while (q + 1 < d + n && !(q[0] == 0xFF && q[1] != 0 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
        i = (size_t)(q - d);
      }
      continue;
    }
    i += L;
  }
  for (int c = 0; c < f->nc; ++c)
    if (!f->qt_ok[f->comp[c].tq]) return fail("jpeg: missing quantisation table");
  return OCRB_OK;
}

// ===================================================================================================================
// host: PNG (inflate + filters + the expansions `png` 0.16 applies for the image crate)
// ===================================================================================================================
struct ZBits {
  const uint8_t *p, *end;
  uint64_t acc = 0;
  int n = 0, pad = 0;  // pad: zero bits appended behind the end of the input
  inline void need(int k) {
    while (n < k) {
      uint64_t b = 0;
      if (p < end) b = *p++; else pad += 8;
      acc |= b << n;
      n += 8;
    }
  }
  inline bool over() const { return pad > n; }  // bits behind the end were CONSUMED (looking ahead is fine)
  inline uint32_t peek(int k) { need(k); return (uint32_t)(acc & ((1ull << k) - 1)); }
  inline void skip(int k) { acc >>= k; n -= k; }
  inline uint32_t get(int k) { if (!k) return 0; const uint32_t v = peek(k); skip(k); return v; }
};

struct ZHuff {
  static constexpr int FAST = 10;
  uint16_t fast[1 << FAST];  // (symbol << 4) | length, 0 = long code
  uint16_t count[16], first_sym[16], sorted[288];
  uint32_t first_code[16];
  bool build(const uint8_t *lens, int n) {
    memset(fast, 0, sizeof fast);
    memset(count, 0, sizeof count);
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    uint32_t code = 0;
    int k = 0;
    uint32_t next[16];
    uint16_t offs[16];
    for (int l = 1; l < 16; ++l) {
      code = (code + count[l - 1]) << 1;
      first_code[l] = next[l] = code;
      first_sym[l] = offs[l] = (uint16_t)k;
      k += count[l];
      if (code + count[l] > (1u << l)) return false;
    }
    for (int i = 0; i < n; ++i) {
      const int l = lens[i];
      if (!l) continue;
      sorted[offs[l]++] = (uint16_t)i;
      const uint32_t c = next[l]++;
      if (l <= FAST) {
        uint32_t rev = 0;
        for (int b = 0; b < l; ++b) rev |= ((c >> b) & 1u) << (l - 1 - b);
        for (uint32_t f = rev; f < (1u << FAST); f += 1u << l) fast[f] = (uint16_t)((i << 4) | l);
      }
    }
    return true;
  }
  inline int decode(ZBits &b) const {
    const uint16_t f = fast[b.peek(FAST)];
    if (f) { b.skip(f & 15); return f >> 4; }
    uint32_t code = 0;
    b.need(15);
    for (int l = 1; l < 16; ++l) {
      code = (code << 1) | (uint32_t)((b.acc >> (l - 1)) & 1);
      if (count[l] && code >= first_code[l] && code - first_code[l] < count[l]) {
        b.skip(l);
        return sorted[first_sym[l] + (code - first_code[l])];
      }
    }
    return -1;
  }
};

bool inflate_zlib(const uint8_t *src, size_t n, std::vector<uint8_t> &out, size_t expect) {
  static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  if (n < 6 || (src[0] & 15) != 8 || ((src[0] << 8) | src[1]) % 31 != 0 || (src[1] & 0x20)) return false;
  ZBits b{src + 2, src + n};
  out.clear();
  out.reserve(expect);
  std::unique_ptr<ZHuff> lit(new ZHuff), dist(new ZHuff);
  for (bool last = false; !last;) {
    last = b.get(1);
    const int type = (int)b.get(2);
    if (type == 0) {
      b.skip(b.n & 7);
      const uint32_t len = b.get(16), nlen = b.get(16);
      if ((len ^ 0xFFFF) != nlen) return false;
      for (uint32_t k = 0; k < len; ++k) out.push_back((uint8_t)b.get(8));
      if (b.over()) return false;
      continue;
    }
    if (type == 3) return false;
    uint8_t lens[320];
    if (type == 1) {
      for (int k = 0; k < 288; ++k) lens[k] = k < 144 ? 8 : k < 256 ? 9 : k < 280 ? 7 : 8;
      lit->build(lens, 288);
      for (int k = 0; k < 30; ++k) lens[k] = 5;
      dist->build(lens, 30);
    } else {
      const int hlit = (int)b.get(5) + 257, hdist = (int)b.get(5) + 1, hclen = (int)b.get(4) + 4;
      if (hlit > 286 || hdist > 30) return false;
      uint8_t cl[19] = {0};
      for (int k = 0; k < hclen; ++k) cl[order[k]] = (uint8_t)b.get(3);
      ZHuff clh;
      if (!clh.build(cl, 19)) return false;
      int k = 0;
      while (k < hlit + hdist) {
        const int s = clh.decode(b);
        if (s < 0 || b.over()) return false;
        if (s < 16) { lens[k++] = (uint8_t)s; continue; }
        int rep, val = 0;
        if (s == 16) { if (k == 0) return false; val = lens[k - 1]; rep = 3 + (int)b.get(2); }
        else if (s == 17) rep = 3 + (int)b.get(3);
        else rep = 11 + (int)b.get(7);
        if (k + rep > hlit + hdist) return false;
        while (rep--) lens[k++] = (uint8_t)val;
      }
      if (!lit->build(lens, hlit) || !dist->build(lens + hlit, hdist)) return false;
    }
    for (;;) {
      const int s = lit->decode(b);
      if (s < 0 || b.over()) return false;
      if (s < 256) { out.push_back((uint8_t)s); continue; }
      if (s == 256) break;
      if (s > 285) return false;
      const size_t len = len_base[s - 257] + b.get(len_extra[s - 257]);
      const int ds = dist->decode(b);
      if (ds < 0 || ds > 29) return false;
      const size_t back = dist_base[ds] + b.get(dist_extra[ds]);
      if (back > out.size()) return false;
      const size_t from = out.size() - back;
      for (size_t k = 0; k < len; ++k) out.push_back(out[from + k]);
    }
  }
  return true;
}

struct PngInfo {
  int W = 0, H = 0, depth = 0, ctype = 0, channels_out = 0;
  bool interlaced = false, has_trns = false;
};

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int png_parse(const uint8_t *d, size_t n, PngInfo *info) {
  if (n < 33 || be32(d + 8) != 13 || memcmp(d + 12, "IHDR", 4) != 0) { set_error("png: no IHDR"); return OCRB_ERR_INVALID; }
  info->W = (int)be32(d + 16);
  info->H = (int)be32(d + 20);
  info->depth = d[24];
  info->ctype = d[25];
  info->interlaced = d[28] != 0;
  if (info->W <= 0 || info->H <= 0) { set_error("png: bad size"); return OCRB_ERR_INVALID; }
  if (info->interlaced || info->depth == 16) { set_error("png: interlaced and 16-bit files are not supported"); return OCRB_ERR_INVALID; }
  const int ct = info->ctype, dp = info->depth;
  const bool ok = (ct == 0 && (dp == 1 || dp == 2 || dp == 4 || dp == 8)) || (ct == 3 && (dp == 1 || dp == 2 || dp == 4 || dp == 8)) ||
                  ((ct == 2 || ct == 4 || ct == 6) && dp == 8);
  if (!ok) { set_error("png: colour type %d with %d bits is not supported", ct, dp); return OCRB_ERR_INVALID; }
  // tRNS adds an alpha channel: look ahead for it
  for (size_t pos = 8; pos + 12 <= n;) {
    const size_t len = be32(d + pos);
    if (pos + 12 + len > n) break;
    if (memcmp(d + pos + 4, "tRNS", 4) == 0) info->has_trns = true;
    if (memcmp(d + pos + 4, "IDAT", 4) == 0) break;
    pos += 12 + len;
  }
  const int base = ct == 0 ? 1 : ct == 2 ? 3 : ct == 3 ? 3 : ct == 4 ? 2 : 4;
  info->channels_out = base + ((info->has_trns && (ct == 0 || ct == 2 || ct == 3)) ? 1 : 0);
  return OCRB_OK;
}

inline int paeth(int a, int b, int c) {
  const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// -> pixels [H][W][channels_out] u8
int png_decode_pixels(const uint8_t *d, size_t n, const PngInfo &info, uint8_t *out, std::string *err) {
  auto fail = [&](const char *what) { *err = what; return OCRB_ERR_INVALID; };
  std::vector<uint8_t> z, raw, plte, trns;
  for (size_t pos = 8; pos + 12 <= n;) {
    const size_t len = be32(d + pos);
    if (pos + 12 + len > n) return fail("png: truncated chunk");
    const uint8_t *body = d + pos + 8;
    if (memcmp(d + pos + 4, "IDAT", 4) == 0) z.insert(z.end(), body, body + len);
    else if (memcmp(d + pos + 4, "PLTE", 4) == 0) plte.assign(body, body + len);
    else if (memcmp(d + pos + 4, "tRNS", 4) == 0) trns.assign(body, body + len);
    else if (memcmp(d + pos + 4, "IEND", 4) == 0) break;
    pos += 12 + len;
  }
  const int ch = info.ctype == 0 ? 1 : info.ctype == 2 ? 3 : info.ctype == 3 ? 1 : info.ctype == 4 ? 2 : 4;
  const int bpp = ch * info.depth / 8 > 1 ? ch * info.depth / 8 : 1;
  const size_t stride = ((size_t)info.W * ch * info.depth + 7) / 8;
  if (!inflate_zlib(z.data(), z.size(), raw, (stride + 1) * info.H)) return fail("png: corrupt compressed data");
  if (raw.size() < (stride + 1) * (size_t)info.H) return fail("png: too little image data");
  if (info.ctype == 3 && plte.size() < 3) return fail("png: palette missing");
  std::vector<uint8_t> zero(stride, 0);
  uint8_t *prev = zero.data();
  for (int y = 0; y < info.H; ++y) {
    uint8_t *cur = raw.data() + (size_t)y * (stride + 1) + 1;
    const int filter = cur[-1];
    if (filter > 4) return fail("png: bad filter type");
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)bpp ? cur[x - bpp] : 0, b = prev[x], c = x >= (size_t)bpp ? prev[x - bpp] : 0;
      const int p = filter == 0 ? 0 : filter == 1 ? a : filter == 2 ? b : filter == 3 ? (a + b) >> 1 : paeth(a, b, c);
      cur[x] = (uint8_t)(cur[x] + p);
    }
    prev = cur;
    // expand this row
    uint8_t *o = out + (size_t)y * info.W * info.channels_out;
    const int maxv = (1 << info.depth) - 1;
    for (int x = 0; x < info.W; ++x) {
      int s[4];
      if (info.depth == 8) {
        for (int k = 0; k < ch; ++k) s[k] = cur[(size_t)x * ch + k];
      } else {
        const int per = 8 / info.depth;
        s[0] = (cur[x / per] >> ((per - 1 - x % per) * info.depth)) & maxv;
      }
      if (info.ctype == 3) {
        const size_t idx = (size_t)s[0];
        const uint8_t *rgb = idx * 3 + 2 < plte.size() ? plte.data() + idx * 3 : plte.data();
        o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2];
        if (info.has_trns) o[3] = idx < trns.size() ? trns[idx] : 255;
      } else if (info.ctype == 0) {
        o[0] = (uint8_t)(info.depth == 8 ? s[0] : s[0] * (255 / maxv));
        if (info.has_trns) o[1] = (trns.size() >= 2 && s[0] == ((trns[0] << 8) | trns[1])) ? 0 : 255;
      } else if (info.ctype == 2) {
        o[0] = (uint8_t)s[0]; o[1] = (uint8_t)s[1]; o[2] = (uint8_t)s[2];
        if (info.has_trns) o[3] = (trns.size() >= 6 && s[0] == trns[1] && s[1] == trns[3] && s[2] == trns[5] && !trns[0] && !trns[2] && !trns[4]) ? 0 : 255;
      } else {
        for (int k = 0; k < ch; ++k) o[k] = (uint8_t)s[k];
      }
      o += info.channels_out;
    }
  }
  return OCRB_OK;
}

}  // namespace

// ===================================================================================================================
// device
// ===================================================================================================================
struct IdctComp {       // one colour component of one JPEG image
  int64_t coef_off;     // int16 elements
  int64_t plane_off;    // bytes
  int bw, bh;
  int qt_off;           // uint16 elements into the quantisation-table array
  int pad;
};

struct AsmImage {
  int kind;             // 0 = interleaved pixels (PNG), 1 = JPEG component planes
  int W, H, nc;
  int64_t out_off;      // bytes into the output arena
  int64_t src_off[3];   // kind 0: [0] = pixel offset; kind 1: plane offsets
  int stride[3], cw[3], chh[3], hs[3], vs[3];  // plane row stride, real size, 2 = upsample that axis
};

#define OCRB_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                             \
  int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                        \
  p2 = s2; p3 = s6;                                                                              \
  p1 = (p2 + p3) * 2217;                                                                         \
  t2 = p1 + p3 * -7567;                                                                          \
  t3 = p1 + p2 * 3135;                                                                           \
  p2 = s0; p3 = s4;                                                                              \
  t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                                  \
  x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                        \
  t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                            \
  p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                                        \
  p5 = (p3 + p4) * 4816;                                                                         \
  t0 = t0 * 1223; t1 = t1 * 8410; t2 = t2 * 12586; t3 = t3 * 6149;                               \
  p1 = p5 + p1 * -3685; p2 = p5 + p2 * -10497; p3 = p3 * -8034; p4 = p4 * -1597;                 \
  t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

__device__ __forceinline__ unsigned clamp255(int v) { return (unsigned)min(max(v, 0), 255); }

constexpr int IDCT_THREADS = 256;  // 32 blocks per CTA
__global__ void __launch_bounds__(IDCT_THREADS) jpeg_idct_kernel(const int16_t *__restrict__ coef, const uint16_t *__restrict__ qts,
                                                                   const IdctComp *__restrict__ comps, uint8_t *__restrict__ planes) {
  __shared__ int tmp[IDCT_THREADS / 8][72];
  const IdctComp c = comps[blockIdx.y];
  const int lb = threadIdx.x >> 3, i = threadIdx.x & 7;
  const int64_t blk = (int64_t)blockIdx.x * (IDCT_THREADS / 8) + lb;
  const bool live = blk < (int64_t)c.bw * c.bh;
  if (live) {
    const int16_t *q = coef + c.coef_off + blk * 64 + i;
    const uint16_t *t = qts + c.qt_off + i;
    const int s0 = q[0] * (int)t[0], s1 = q[8] * (int)t[8], s2 = q[16] * (int)t[16], s3 = q[24] * (int)t[24];
    const int s4 = q[32] * (int)t[32], s5 = q[40] * (int)t[40], s6 = q[48] * (int)t[48], s7 = q[56] * (int)t[56];
    OCRB_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)
    x0 += 512; x1 += 512; x2 += 512; x3 += 512;
    int *w = tmp[lb] + i;
    w[0] = (x0 + t3) >> 10; w[63] = (x0 - t3) >> 10;
    w[9] = (x1 + t2) >> 10; w[54] = (x1 - t2) >> 10;
    w[18] = (x2 + t1) >> 10; w[45] = (x2 - t1) >> 10;
    w[27] = (x3 + t0) >> 10; w[36] = (x3 - t0) >> 10;
  }
  __syncwarp();
  if (live) {
    const int *s = tmp[lb] + i * 9;
    OCRB_IDCT_1D(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7])
    const int bias = 65536 + (128 << 17);
    x0 += bias; x1 += bias; x2 += bias; x3 += bias;
    uint2 o;
    o.x = clamp255((x0 + t3) >> 17) | (clamp255((x1 + t2) >> 17) << 8) | (clamp255((x2 + t1) >> 17) << 16) | (clamp255((x3 + t0) >> 17) << 24);
    o.y = clamp255((x3 - t0) >> 17) | (clamp255((x2 - t1) >> 17) << 8) | (clamp255((x1 - t2) >> 17) << 16) | (clamp255((x0 - t3) >> 17) << 24);
    const int by = (int)(blk / c.bw), bx = (int)(blk % c.bw);
    *reinterpret_cast<uint2 *>(planes + c.plane_off + ((int64_t)(by * 8 + i) * c.bw + bx) * 8) = o;
  }
}

// one chroma (or luma) sample of output pixel (x, y): upsampler.rs in closed form
__device__ __forceinline__ int plane_sample(const uint8_t *__restrict__ pl, int stride, int cw, int chh, int hs, int vs, int x, int y) {
  if (hs == 1 && vs == 1) return pl[(int64_t)y * stride + x];
  if (vs == 1) {  // H2V1
    const uint8_t *r = pl + (int64_t)y * stride;
    const int i = x >> 1;
    if (cw == 1 || x == 0) return r[0];
    if (x == 2 * cw - 1) return r[cw - 1];
    return (3 * r[i] + r[(x & 1) ? i + 1 : i - 1] + 2) >> 2;
  }
  const int k = y >> 1;
  const int kf = (y & 1) ? min(k + 1, chh - 1) : max(k - 1, 0);
  const uint8_t *nr = pl + (int64_t)k * stride, *fr = pl + (int64_t)kf * stride;
  if (hs == 1) return (3 * nr[x] + fr[x] + 2) >> 2;  // H1V2
  const int i = x >> 1;
  const int ti = 3 * nr[i] + fr[i];
  if (cw == 1 || x == 0 || x == 2 * cw - 1) return (ti + 2) >> 2;
  const int j = (x & 1) ? i + 1 : i - 1;
  return (3 * ti + 3 * nr[j] + fr[j] + 8) >> 4;
}

// format: 0 = RGBA8 (into_rgba), 1 = luma (into_luma)
__global__ void decode_assemble_kernel(const uint8_t *__restrict__ planes, const uint8_t *__restrict__ pixels, const AsmImage *__restrict__ imgs,
                                       int format, uint8_t *__restrict__ out) {
  const AsmImage &im = imgs[blockIdx.y];
  const int64_t total = (int64_t)im.W * im.H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % im.W), y = (int)(idx / im.W);
    int r, g, b, a = 255, grey = -1;
    if (im.kind == 0) {
      const uint8_t *p = pixels + im.src_off[0] + idx * im.nc;
      if (im.nc <= 2) {
        grey = p[0];
        r = g = b = grey;
        if (im.nc == 2) a = p[1];
      } else {
        r = p[0]; g = p[1]; b = p[2];
        if (im.nc == 4) a = p[3];
      }
    } else if (im.nc == 1) {
      grey = planes[im.src_off[0] + (int64_t)y * im.stride[0] + x];
      r = g = b = grey;
    } else {
      const float Y = (float)plane_sample(planes + im.src_off[0], im.stride[0], im.cw[0], im.chh[0], im.hs[0], im.vs[0], x, y);
      const float cb = (float)plane_sample(planes + im.src_off[1], im.stride[1], im.cw[1], im.chh[1], im.hs[1], im.vs[1], x, y) - 128.0f;
      const float cr = (float)plane_sample(planes + im.src_off[2], im.stride[2], im.cw[2], im.chh[2], im.hs[2], im.vs[2], x, y) - 128.0f;
      // decoder.rs ycbcr_to_rgb (0.1.20): f32, +0.5, truncating cast, clamp (-fmad=false: each product rounds)
      const float rf = Y + 1.40200f * cr;
      const float gf = Y - 0.34414f * cb - 0.71414f * cr;
      const float bf = Y + 1.77200f * cb;
      r = min(max((int)(rf + 0.5f), 0), 255);
      g = min(max((int)(gf + 0.5f), 0), 255);
      b = min(max((int)(bf + 0.5f), 0), 255);
    }
    if (format == 0) {
      uchar4 o;
      o.x = (uint8_t)r; o.y = (uint8_t)g; o.z = (uint8_t)b; o.w = (uint8_t)a;
      *reinterpret_cast<uchar4 *>(out + im.out_off + idx * 4) = o;
    } else {
      // image 0.23.11 into_luma: grey sources pass through, RGB -> Rec.709 weights in f32, truncating cast
      const float l = 0.2126f * (float)r + 0.7152f * (float)g + 0.0722f * (float)b;
      out[im.out_off + idx] = grey >= 0 ? (uint8_t)grey : (uint8_t)l;
    }
  }
}

// ===================================================================================================================
// batch driver
// ===================================================================================================================
namespace {

enum { FMT_JPEG = 1, FMT_PNG = 2 };

struct FileInfo {
  int fmt = 0, W = 0, H = 0, channels = 0;
  JFrame jf;
  PngInfo png;
};

int sniff(const uint8_t *d, size_t n, FileInfo *fi) {
  if (d && n >= 4 && d[0] == 0xFF && d[1] == 0xD8) {
    fi->fmt = FMT_JPEG;
    OCRB_TRY(jpeg_parse_frame(d, n, &fi->jf));
    fi->W = fi->jf.W;
    fi->H = fi->jf.H;
    fi->channels = fi->jf.nc;
    return OCRB_OK;
  }
  if (d && n >= 8 && memcmp(d, "\x89PNG\r\n\x1a\n", 8) == 0) {
    fi->fmt = FMT_PNG;
    OCRB_TRY(png_parse(d, n, &fi->png));
    fi->W = fi->png.W;
    fi->H = fi->png.H;
    fi->channels = fi->png.channels_out;
    return OCRB_OK;
  }
  set_error("unsupported image format (JPEG and PNG are decoded)");  // image::open -> Err
  return OCRB_ERR_INVALID;
}

}  // namespace

// Decodes n files into `out_dev` (device): image i at byte offset out_offsets[i], RGBA8 (format 0) or luma (1).
// `infos` comes from sniff().
static int decode_batch_device(ocrb_ctx *ctx, const uint8_t *const *files, const size_t *sizes, int n, std::vector<FileInfo> &infos, int format,
                               const int64_t *out_offsets, uint8_t *out_dev) {
  // arena layout
  int64_t coef_elems = 0, plane_bytes = 0, pixel_bytes = 0;
  int n_comps = 0;
  std::vector<int64_t> pix_off((size_t)n, 0);
  for (int i = 0; i < n; ++i) {
    FileInfo &fi = infos[i];
    if (fi.fmt == FMT_JPEG) {
      for (int c = 0; c < fi.jf.nc; ++c) {
        JComp &k = fi.jf.comp[c];
        k.coef_off = coef_elems;
        k.plane_off = plane_bytes;
        coef_elems += (int64_t)k.bw * k.bh * 64;
        plane_bytes += (int64_t)k.bw * k.bh * 64;
        ++n_comps;
      }
    } else {
      pix_off[i] = pixel_bytes;
      pixel_bytes += ((int64_t)fi.W * fi.H * fi.channels + 15) / 16 * 16;
    }
  }
  // pinned staging: [coefficients | PNG pixels | quantisation tables | IdctComp[] | AsmImage[]]
  const size_t o_pix = (size_t)coef_elems * 2, o_qt = o_pix + (size_t)pixel_bytes, o_ic = o_qt + (size_t)n * 4 * 64 * 2;
  const size_t o_ai = (o_ic + (size_t)n_comps * sizeof(IdctComp) + 15) / 16 * 16, total = o_ai + (size_t)n * sizeof(AsmImage);
  OCRB_TRY(ctx->pin[2].reserve(total));
  uint8_t *host = ctx->pin[2].as<uint8_t>();
  memset(host, 0, o_pix);  // coefficients start at zero (progressive scans accumulate)
  int16_t *coef_h = reinterpret_cast<int16_t *>(host);
  // host workers: one image at a time per thread
  std::atomic<int> next{0}, failed{-1};
  std::vector<std::string> errs((size_t)n);
  auto work = [&]() {
    for (int i; (i = next.fetch_add(1)) < n;) {
      FileInfo &fi = infos[i];
      int rc = OCRB_ERR_INTERNAL;
      try {
        rc = fi.fmt == FMT_JPEG ? jpeg_decode_coefficients(files[i], sizes[i], &fi.jf, coef_h, &errs[i])
                                : png_decode_pixels(files[i], sizes[i], fi.png, host + o_pix + pix_off[i], &errs[i]);
      } catch (const std::exception &e) {
        errs[i] = e.what();
      }
      if (rc != OCRB_OK) {
        int expect = -1;
        failed.compare_exchange_strong(expect, i);
      }
    }
  };
  unsigned hw = std::thread::hardware_concurrency();
  const int n_threads = (int)std::min<unsigned>(hw ? hw : 4u, (unsigned)n);
  if (n_threads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }
  if (failed.load() >= 0) {
    set_error("image %d: %s", failed.load(), errs[(size_t)failed.load()].c_str());
    return OCRB_ERR_INVALID;
  }
  // descriptors
  uint16_t *qt_h = reinterpret_cast<uint16_t *>(host + o_qt);
  IdctComp *ic = reinterpret_cast<IdctComp *>(host + o_ic);
  AsmImage *ai = reinterpret_cast<AsmImage *>(host + o_ai);
  int64_t max_blocks = 0, max_pixels = 0;
  for (int i = 0, k = 0; i < n; ++i) {
    const FileInfo &fi = infos[i];
    AsmImage &a = ai[i];
    memset(&a, 0, sizeof a);
    a.W = fi.W;
    a.H = fi.H;
    a.nc = fi.channels;
    a.out_off = out_offsets[i];
    max_pixels = std::max<int64_t>(max_pixels, (int64_t)fi.W * fi.H);
    if (fi.fmt == FMT_PNG) {
      a.kind = 0;
      a.src_off[0] = pix_off[i];
      continue;
    }
    a.kind = 1;
    memcpy(qt_h + (size_t)i * 256, fi.jf.qt, sizeof fi.jf.qt);
    for (int c = 0; c < fi.jf.nc; ++c, ++k) {
      const JComp &jc = fi.jf.comp[c];
      ic[k].coef_off = jc.coef_off;
      ic[k].plane_off = jc.plane_off;
      ic[k].bw = jc.bw;
      ic[k].bh = jc.bh;
      ic[k].qt_off = i * 256 + jc.tq * 64;
      ic[k].pad = 0;
      max_blocks = std::max<int64_t>(max_blocks, (int64_t)jc.bw * jc.bh);
      a.src_off[c] = jc.plane_off;
      a.stride[c] = jc.bw * 8;
      a.cw[c] = jc.w;
      a.chh[c] = jc.hpx;
      a.hs[c] = (jc.h == fi.jf.hmax || fi.W == 1) ? 1 : 2;  // upsampler.rs choose_upsampler
      a.vs[c] = (jc.v == fi.jf.vmax || fi.H == 1) ? 1 : 2;
    }
  }
  // device: staging copy (one H2D), planes, two launches
  OCRB_TRY(ctx->stage[5].reserve(total));
  OCRB_TRY(ctx->stage[2].reserve((size_t)plane_bytes + 16));
  uint8_t *dev = ctx->stage[5].as<uint8_t>();
  OCRB_CUDA(cudaMemcpyAsync(dev, host, total, cudaMemcpyHostToDevice, ctx->stream));
  if (n_comps > 0) {
    dim3 grid((unsigned)cdiv(max_blocks, IDCT_THREADS / 8), (unsigned)n_comps);
    jpeg_idct_kernel<<<grid, IDCT_THREADS, 0, ctx->stream>>>(reinterpret_cast<const int16_t *>(dev), reinterpret_cast<const uint16_t *>(dev + o_qt),
                                                            reinterpret_cast<const IdctComp *>(dev + o_ic), ctx->stage[2].as<uint8_t>());
    OCRB_TRY(check_launch(ctx, "jpeg_idct"));
  }
  int64_t bx = cdiv(max_pixels, 256);
  const int64_t cap = std::max<int64_t>(1, (int64_t)ctx->sm_count * 16 / n);
  if (bx > cap) bx = cap;
  decode_assemble_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, ctx->stream>>>(ctx->stage[2].as<uint8_t>(), dev + o_pix,
                                                                                   reinterpret_cast<const AsmImage *>(dev + o_ai), format, out_dev);
  return check_launch(ctx, "decode_assemble");
}

}  // namespace ocrb

using namespace ocrb;

extern "C" {

int ocrb_image_info(const uint8_t *file, size_t size, int *width, int *height, int *channels) {
  OCRB_REQUIRE(file && width && height, "null argument");
  FileInfo fi;
  OCRB_TRY(sniff(file, size, &fi));
  *width = fi.W;
  *height = fi.H;
  if (channels) *channels = fi.channels;
  return OCRB_OK;
}

// host-only test hook: the host stage's product for one file — JPEG: int16 coefficients (components concatenated,
// [block row][block][64], natural order); PNG: pixels [h][w][channels].  *needed = bytes; out may be NULL to size.
int ocrb_debug_decode_host(const uint8_t *file, size_t size, void *out, size_t cap, size_t *needed) {
  OCRB_REQUIRE(file && needed, "null argument");
  try {
    FileInfo fi;
    OCRB_TRY(sniff(file, size, &fi));
    size_t bytes = 0;
    if (fi.fmt == FMT_JPEG) {
      for (int c = 0; c < fi.jf.nc; ++c) {
        fi.jf.comp[c].coef_off = (int64_t)(bytes / 2);
        bytes += (size_t)fi.jf.comp[c].bw * fi.jf.comp[c].bh * 128;
      }
    } else {
      bytes = (size_t)fi.W * fi.H * fi.channels;
    }
    *needed = bytes;
    if (!out) return OCRB_OK;
    if (cap < bytes) { set_error("output buffer too small"); return OCRB_ERR_CAPACITY; }
    std::string err;
    memset(out, 0, bytes);
    const int rc = fi.fmt == FMT_JPEG ? jpeg_decode_coefficients(file, size, &fi.jf, (int16_t *)out, &err) : png_decode_pixels(file, size, fi.png, (uint8_t *)out, &err);
    if (rc != OCRB_OK) set_error("%s", err.c_str());
    return rc;
  } catch (const std::exception &e) {
    set_error("decode: %s", e.what());
    return OCRB_ERR_INTERNAL;
  }
}

int ocrb_decode_images(ocrb_ctx *ctx, const uint8_t *const *files, const size_t *sizes, int n, int format, const int64_t *out_offsets, uint8_t *out) {
  OCRB_REQUIRE(ctx && files && sizes && out_offsets && out && n > 0, "bad argument");
  OCRB_REQUIRE(format == OCRB_PIXELS_RGBA || format == OCRB_PIXELS_LUMA, "format must be OCRB_PIXELS_RGBA or OCRB_PIXELS_LUMA");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  try {
    std::vector<FileInfo> infos((size_t)n);
    const int bpp = format == OCRB_PIXELS_RGBA ? 4 : 1;
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
      OCRB_TRY(sniff(files[i], sizes[i], &infos[i]));
      OCRB_REQUIRE(out_offsets[i] >= 0 && out_offsets[i] % bpp == 0, "image %d: bad output offset", i);
      total = std::max<int64_t>(total, out_offsets[i] + (int64_t)infos[i].W * infos[i].H * bpp);
    }
    void *dst = nullptr;
    OCRB_TRY(out_device(ctx, 1, out, (size_t)total, &dst));
    OCRB_TRY(decode_batch_device(ctx, files, sizes, n, infos, format, out_offsets, (uint8_t *)dst));
    OCRB_TRY(finish_output(ctx, out, dst, (size_t)total));
    return sync(ctx);
  } catch (const std::exception &e) {
    set_error("decode: %s", e.what());
    return OCRB_ERR_INTERNAL;
  }
}

int ocrb_preprocess_files(ocrb_ctx *ctx, const uint8_t *const *files, const size_t *sizes, int n, int W, int H, uint8_t *out_gray, double *adjust) {
  OCRB_REQUIRE(ctx && files && sizes && out_gray && adjust && n > 0 && W > 0 && H > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  try {
    std::vector<FileInfo> infos((size_t)n);
    std::vector<int64_t> offs((size_t)n);
    std::vector<int> ws((size_t)n), hs((size_t)n);
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
      OCRB_TRY(sniff(files[i], sizes[i], &infos[i]));
      offs[i] = total;
      ws[i] = infos[i].W;
      hs[i] = infos[i].H;
      total += (int64_t)infos[i].W * infos[i].H * 4;
    }
    // the decoded RGBA arena stays on the device: it is what the fused resize / luma / pad kernel reads
    DevBuf &arena = ctx->decode_rgba;
    OCRB_TRY(arena.reserve((size_t)total));
    OCRB_TRY(decode_batch_device(ctx, files, sizes, n, infos, OCRB_PIXELS_RGBA, offs.data(), arena.as<uint8_t>()));
    return ocrb_preprocess_rgba_batch(ctx, arena.as<uint8_t>(), offs.data(), ws.data(), hs.data(), n, W, H, out_gray, adjust);
  } catch (const std::exception &e) {
    set_error("decode: %s", e.what());
    return OCRB_ERR_INTERNAL;
  }
}

}  // extern "C"

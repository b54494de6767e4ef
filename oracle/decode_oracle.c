// TEST INFRASTRUCTURE — CPU oracle of the file decode in front of image_ops::preprocess_image
// (image_ops.rs:193 `open(file)?.into_rgba()`).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
// use this file; nothing under ocr_rs_b200/ does.
//
// The reference decodes through `image` 0.23.11, which hands JPEG files to `jpeg-decoder` 0.1.20 (Cargo.lock:667).
// That crate's sources are absent here, so its published algorithm is restated:
//   * entropy decoding: ITU-T T.81 Annex F (sequential) and Annex G (progressive) — the coefficients are fixed by the
//     standard, any conforming decoder produces the same ones;
//   * dequantisation + inverse DCT: jpeg-decoder's idct.rs is a port of stb_image's integer IDCT
//     (12-bit constants, column pass >> 10 with +512, row pass >> 17 with +65536 + (128 << 17), clamp to u8);
//   * chroma upsampling: upsampler.rs — H2V1 / H1V2 / H2V2 "triangle" filters on the component's REAL size
//     (ceil(image size * factor / max factor)), H1V1 copy;
//   * colour conversion: decoder.rs ycbcr_to_rgb.  Two forms existed in the 0.1.x line (f32 arithmetic with +0.5
//     truncation, and a 20-bit fixed-point form); `variant` selects one, and tests/test_decode_oracle.py shows
//     which one reproduces the reference's fixtures.
// PINNED: the reference's own test (image_ops.rs:805-1008) asserts preprocess_image(img{55,224,494,545}.jpg) ==
// test_data/preprocessed_img*.png; decode (this file) + resize/luma/pad (postproc_oracle.c) reproduces all four
// fixtures bit for bit (three baseline files, one progressive; 4:2:0 and 4:4:4).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint8_t bits[17];
  uint8_t vals[256];
  int mincode[17], maxcode[18], valptr[17];
  int present;
} HuffTable;

typedef struct {
  int id, h, v, tq;
  int w, h_px;          // real size of the component plane (ceil(image * factor / max))
  int bw, bh;           // block grid of the plane buffer (MCU-padded)
  int16_t *coef;        // [bh][bw][64], natural order
  uint8_t *plane;       // [bh * 8][bw * 8]
  int dc_pred;
  int td, ta;
} Comp;

typedef struct {
  const uint8_t *p, *end;
  uint32_t bitbuf;
  int bitcnt;
  int marker;           // pending marker met inside the entropy-coded segment (0 = none)
} Bits;

static const uint8_t ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                   15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static void huff_build(HuffTable *t) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t->valptr[l] = k;
    t->mincode[l] = code;
    code += t->bits[l];
    k += t->bits[l];
    t->maxcode[l] = t->bits[l] ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[17] = 0x7fffffff;
  t->present = 1;
}

static int bits_fill(Bits *b) {  // one more byte into the buffer; past a marker the stream reads as zeros
  int byte = 0;
  if (!b->marker && b->p < b->end) {
    byte = *b->p++;
    if (byte == 0xFF) {
      int nx = b->p < b->end ? *b->p : 0xD9;
      while (nx == 0xFF && b->p + 1 < b->end) { ++b->p; nx = *b->p; }  // fill bytes
      if (nx == 0) {
        ++b->p;
      } else {
        b->marker = nx;
        ++b->p;
        byte = 0;
      }
    }
  }
  b->bitbuf = (b->bitbuf << 8) | (uint32_t)byte;
  b->bitcnt += 8;
  return 0;
}
static inline int bits_get(Bits *b, int n) {
  if (n == 0) return 0;
  while (b->bitcnt < n) bits_fill(b);
  b->bitcnt -= n;
  return (int)((b->bitbuf >> b->bitcnt) & ((1u << n) - 1));
}
static inline int huff_decode(Bits *b, const HuffTable *t) {
  int code = 0;
  for (int l = 1; l <= 16; ++l) {
    code = (code << 1) | bits_get(b, 1);
    if (t->maxcode[l] >= 0 && code <= t->maxcode[l] && code >= t->mincode[l]) return t->vals[t->valptr[l] + code - t->mincode[l]];
  }
  return -1;
}
static inline int extend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// ---- IDCT (jpeg-decoder idct.rs = stb_image stbi__idct_block) ----------------------------------------------------
static int f2f(float x) { return (int)(x * 4096.0f + 0.5f); }
#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                   \
  int32_t t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                     \
  p2 = s2; p3 = s6;                                                                               \
  p1 = (p2 + p3) * K[0];                                                                          \
  t2 = p1 + p3 * K[1];                                                                            \
  t3 = p1 + p2 * K[2];                                                                            \
  p2 = s0; p3 = s4;                                                                               \
  t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                                   \
  x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                         \
  t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                             \
  p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                                         \
  p5 = (p3 + p4) * K[3];                                                                          \
  t0 = t0 * K[4]; t1 = t1 * K[5]; t2 = t2 * K[6]; t3 = t3 * K[7];                                 \
  p1 = p5 + p1 * K[8]; p2 = p5 + p2 * K[9]; p3 = p3 * K[10]; p4 = p4 * K[11];                     \
  t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

static inline uint8_t clamp_u8(int32_t v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

static void idct_block(const int16_t *c, const uint16_t *q, uint8_t *out, int stride) {
  int32_t K[12] = {f2f(0.5411961f), f2f(-1.847759065f), f2f(0.765366865f), f2f(1.175875602f), f2f(0.298631336f), f2f(2.053119869f),
                   f2f(3.072711026f), f2f(1.501321110f), f2f(-0.899976223f), f2f(-2.562915447f), f2f(-1.961570560f), f2f(-0.390180644f)};
  int32_t tmp[64];
  for (int i = 0; i < 8; ++i) {
    int32_t s0 = c[i] * q[i], s1 = c[i + 8] * q[i + 8], s2 = c[i + 16] * q[i + 16], s3 = c[i + 24] * q[i + 24];
    int32_t s4 = c[i + 32] * q[i + 32], s5 = c[i + 40] * q[i + 40], s6 = c[i + 48] * q[i + 48], s7 = c[i + 56] * q[i + 56];
    IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)
    x0 += 512; x1 += 512; x2 += 512; x3 += 512;
    tmp[i] = (x0 + t3) >> 10; tmp[i + 56] = (x0 - t3) >> 10;
    tmp[i + 8] = (x1 + t2) >> 10; tmp[i + 48] = (x1 - t2) >> 10;
    tmp[i + 16] = (x2 + t1) >> 10; tmp[i + 40] = (x2 - t1) >> 10;
    tmp[i + 24] = (x3 + t0) >> 10; tmp[i + 32] = (x3 - t0) >> 10;
  }
  for (int i = 0; i < 8; ++i) {
    const int32_t *s = tmp + i * 8;
    IDCT_1D(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7])
    x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
    uint8_t *o = out + i * stride;
    o[0] = clamp_u8((x0 + t3) >> 17); o[7] = clamp_u8((x0 - t3) >> 17);
    o[1] = clamp_u8((x1 + t2) >> 17); o[6] = clamp_u8((x1 - t2) >> 17);
    o[2] = clamp_u8((x2 + t1) >> 17); o[5] = clamp_u8((x2 - t1) >> 17);
    o[3] = clamp_u8((x3 + t0) >> 17); o[4] = clamp_u8((x3 - t0) >> 17);
  }
}

// ---- upsampling (jpeg-decoder upsampler.rs) ----------------------------------------------------------------------
static void upsample_row(const Comp *c, int hmax, int vmax, int out_w, int out_h, int row, uint8_t *out) {
  const int h1 = c->h == hmax || out_w == 1, v1 = c->v == vmax || out_h == 1;
  const int stride = c->bw * 8, iw = c->w, ih = c->h_px;
  const uint8_t *in = c->plane;
  if (h1 && v1) {
    memcpy(out, in + (size_t)row * stride, (size_t)out_w);
    return;
  }
  if (!h1 && v1) {  // H2V1
    const uint8_t *r = in + (size_t)row * stride;
    if (iw == 1) { out[0] = out[1] = r[0]; return; }
    out[0] = r[0];
    out[1] = (uint8_t)((r[0] * 3u + r[1] + 2) >> 2);
    for (int i = 1; i < iw - 1; ++i) {
      unsigned s = 3u * r[i] + 2;
      out[2 * i] = (uint8_t)((s + r[i - 1]) >> 2);
      out[2 * i + 1] = (uint8_t)((s + r[i + 1]) >> 2);
    }
    out[(iw - 1) * 2] = (uint8_t)((r[iw - 1] * 3u + r[iw - 2] + 2) >> 2);
    out[(iw - 1) * 2 + 1] = r[iw - 1];
    return;
  }
  const float row_near = (float)row / 2.0f;
  const float fract = row_near - (float)(int)row_near;
  float row_far = row_near + fract * 3.0f - 0.25f;
  if (row_far > (float)(ih - 1)) row_far = (float)(ih - 1);
  const uint8_t *near = in + (size_t)(int)row_near * stride;
  const uint8_t *far = in + (size_t)(row_far < 0 ? 0 : (int)row_far) * stride;
  if (h1) {  // H1V2
    for (int i = 0; i < iw; ++i) out[i] = (uint8_t)((3u * near[i] + far[i] + 2) >> 2);
    return;
  }
  // H2V2
  if (iw == 1) { out[0] = out[1] = (uint8_t)((3u * near[0] + far[0] + 2) >> 2); return; }
  unsigned t0 = 3u * near[0] + far[0], t1 = 3u * near[1] + far[1];
  out[0] = (uint8_t)((t0 + 2) >> 2);
  out[1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
  for (int i = 2; i < iw; ++i) {
    unsigned t2 = 3u * near[i] + far[i];
    out[i * 2 - 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
    out[i * 2 - 1] = (uint8_t)((3 * t1 + t2 + 8) >> 4);
    t0 = t1; t1 = t2;
  }
  out[iw * 2 - 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
  out[iw * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
}

// ---- colour conversion (jpeg-decoder decoder.rs) -----------------------------------------------------------------
static inline uint8_t clampi(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
static void ycbcr_to_rgb(int variant, uint8_t y8, uint8_t cb8, uint8_t cr8, uint8_t *rgb) {
  if (variant == 0) {  // f32 form
    const float y = (float)y8, cb = (float)cb8 - 128.0f, cr = (float)cr8 - 128.0f;  // -ffp-contract=off: no FMA
    const float r = y + 1.40200f * cr;
    const float g = y - 0.34414f * cb - 0.71414f * cr;
    const float b = y + 1.77200f * cb;
    rgb[0] = clampi((int)(r + 0.5f)); rgb[1] = clampi((int)(g + 0.5f)); rgb[2] = clampi((int)(b + 0.5f));
  } else {  // 20-bit fixed point (libjpeg-turbo jdcolext.c constants)
    const int SH = 20, HALF = (1 << SH) / 2;
    const int c1 = (int)(1.40200f * (float)(1 << SH) + 0.5f), c2 = (int)(0.34414f * (float)(1 << SH) + 0.5f);
    const int c3 = (int)(0.71414f * (float)(1 << SH) + 0.5f), c4 = (int)(1.77200f * (float)(1 << SH) + 0.5f);
    const int y = (int)y8 * (1 << SH) + HALF, cb = (int)cb8 - 128, cr = (int)cr8 - 128;
    rgb[0] = clampi((y + c1 * cr) >> SH);
    rgb[1] = clampi((y - c2 * cb - c3 * cr) >> SH);
    rgb[2] = clampi((y + c4 * cb) >> SH);
  }
}

// ---- scan decoding -----------------------------------------------------------------------------------------------
typedef struct {
  int ss, se, ah, al, progressive;
  int eobrun;
} Scan;

static int decode_block_baseline(Bits *b, Comp *c, const HuffTable *dc, const HuffTable *ac, int16_t *coef) {
  int s = huff_decode(b, dc);
  if (s < 0 || s > 11) return -1;
  c->dc_pred += extend(bits_get(b, s), s);
  coef[0] = (int16_t)c->dc_pred;
  for (int k = 1; k < 64;) {
    int rs = huff_decode(b, ac);
    if (rs < 0) return -1;
    int r = rs >> 4;
    s = rs & 15;
    if (s == 0) {
      if (r == 15) { k += 16; continue; }
      break;
    }
    k += r;
    if (k > 63) return -1;
    coef[ZIGZAG[k]] = (int16_t)extend(bits_get(b, s), s);
    ++k;
  }
  return 0;
}

static int decode_block_progressive(Bits *b, Comp *c, const HuffTable *dc, const HuffTable *ac, int16_t *coef, Scan *sc) {
  if (sc->ss == 0) {  // DC scan
    if (sc->ah == 0) {
      int s = huff_decode(b, dc);
      if (s < 0 || s > 11) return -1;
      c->dc_pred += extend(bits_get(b, s), s);
      coef[0] = (int16_t)(c->dc_pred * (1 << sc->al));
    } else if (bits_get(b, 1)) {
      coef[0] |= (int16_t)(1 << sc->al);
    }
    return 0;
  }
  if (sc->ah == 0) {  // AC first pass
    if (sc->eobrun > 0) { --sc->eobrun; return 0; }
    for (int k = sc->ss; k <= sc->se;) {
      int rs = huff_decode(b, ac);
      if (rs < 0) return -1;
      int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r < 15) {
          sc->eobrun = (1 << r) - 1;
          if (r) sc->eobrun += bits_get(b, r);
          break;
        }
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) return -1;
      coef[ZIGZAG[k]] = (int16_t)(extend(bits_get(b, s), s) * (1 << sc->al));
      ++k;
    }
    return 0;
  }
  // AC refinement (T.81 G.1.2.3)
  const int p1 = 1 << sc->al, m1 = -1 * (1 << sc->al);
  int k = sc->ss;
  if (sc->eobrun == 0) {
    for (; k <= sc->se;) {
      int rs = huff_decode(b, ac);
      if (rs < 0) return -1;
      int r = rs >> 4, s = rs & 15, value = 0;
      if (s == 0) {
        if (r < 15) {
          sc->eobrun = (1 << r);
          if (r) sc->eobrun += bits_get(b, r);
          break;
        }
      } else {
        if (s != 1) return -1;
        value = bits_get(b, 1) ? p1 : m1;
      }
      // skip r zero-history coefficients, refining the non-zero ones met on the way
      for (; k <= sc->se; ++k) {
        int16_t *co = &coef[ZIGZAG[k]];
        if (*co != 0) {
          if (bits_get(b, 1) && (*co & p1) == 0) *co = (int16_t)(*co >= 0 ? *co + p1 : *co + m1);
        } else {
          if (r == 0) {
            if (value) *co = (int16_t)value;
            ++k;
            break;
          }
          --r;
        }
      }
    }
  }
  if (sc->eobrun > 0) {
    for (; k <= sc->se; ++k) {
      int16_t *co = &coef[ZIGZAG[k]];
      if (*co != 0 && bits_get(b, 1) && (*co & p1) == 0) *co = (int16_t)(*co >= 0 ? *co + p1 : *co + m1);
    }
    --sc->eobrun;
  }
  return 0;
}

#ifdef DBG
#include <stdio.h>
#define ERR(code) do { fprintf(stderr, "ERR line %d\n", __LINE__); rc = (code); goto done; } while (0)
#else
#define ERR(code) do { rc = (code); goto done; } while (0)
#endif

// Decodes a JFIF / JPEG file into interleaved pixels: n_comp = 1 (L8) or 3 (RGB8).  Returns 0, or a negative code
// (-1 malformed, -2 unsupported: arithmetic coding, lossless, 12-bit, CMYK, sampling ratios other than 1 and 2).
// `coef_out` (optional): the entropy-decoded coefficients of all components, [component][block row][block][64] in
// natural order over the MCU-padded block grid (what the product's host stage hands to its device kernels).
int orc_jpeg_decode_ex(const uint8_t *data, size_t n, int variant, uint8_t **pixels, int *width, int *height, int *n_comp,
                       int16_t **coef_out, size_t *coef_count) {
  int rc = 0;
  uint16_t qt[4][64];
  int qt_present[4] = {0, 0, 0, 0};
  HuffTable hdc[4], hac[4];
  Comp comp[4];
  int nc = 0, W = 0, H = 0, progressive = 0, restart = 0, hmax = 1, vmax = 1, have_sof = 0, adobe_transform = -1;
  uint8_t *line[4] = {0, 0, 0, 0};
  uint8_t *out = NULL;
  memset(hdc, 0, sizeof hdc);
  memset(hac, 0, sizeof hac);
  memset(comp, 0, sizeof comp);
  if (n < 4 || data[0] != 0xFF || data[1] != 0xD8) return -1;
  size_t i = 2;
  for (;;) {
    if (i + 2 > n) ERR(-1);
    if (data[i] != 0xFF) ERR(-1);
    while (i < n && data[i] == 0xFF) ++i;  // marker prefix + fill bytes
    if (i >= n) ERR(-1);
    const int m = data[i++];
    if (m == 0xD9) break;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (i + 2 > n) ERR(-1);
    const size_t L = ((size_t)data[i] << 8) | data[i + 1];
    if (L < 2 || i + L > n) ERR(-1);
    const uint8_t *seg = data + i + 2;
    const size_t sl = L - 2;
    if (m == 0xDB) {
      for (size_t o = 0; o < sl;) {
        const int pq = seg[o] >> 4, tq = seg[o] & 15;
        if (tq > 3 || pq > 1 || o + 1 + 64 * (size_t)(pq + 1) > sl) ERR(-1);
        for (int k = 0; k < 64; ++k) qt[tq][ZIGZAG[k]] = pq ? (uint16_t)((seg[o + 1 + 2 * k] << 8) | seg[o + 2 + 2 * k]) : seg[o + 1 + k];
        qt_present[tq] = 1;
        o += 1 + 64 * (size_t)(pq + 1);
      }
    } else if (m == 0xC4) {
      for (size_t o = 0; o < sl;) {
        if (o + 17 > sl) ERR(-1);
        const int tc = seg[o] >> 4, th = seg[o] & 15;
        if (tc > 1 || th > 3) ERR(-1);
        HuffTable *t = tc ? &hac[th] : &hdc[th];
        int total = 0;
        t->bits[0] = 0;
        for (int l = 1; l <= 16; ++l) { t->bits[l] = seg[o + l]; total += seg[o + l]; }
        if (total > 256 || o + 17 + (size_t)total > sl) ERR(-1);
        memcpy(t->vals, seg + o + 17, (size_t)total);
        huff_build(t);
        o += 17 + (size_t)total;
      }
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
      if (have_sof || sl < 6) ERR(-1);
      if (seg[0] != 8) ERR(-2);
      H = (seg[1] << 8) | seg[2];
      W = (seg[3] << 8) | seg[4];
      nc = seg[5];
      progressive = m == 0xC2;
      if (W == 0 || H == 0) ERR(-1);
      if (nc != 1 && nc != 3) ERR(-2);
      if (sl < 6 + 3 * (size_t)nc) ERR(-1);
      for (int c = 0; c < nc; ++c) {
        comp[c].id = seg[6 + 3 * c];
        comp[c].h = seg[7 + 3 * c] >> 4;
        comp[c].v = seg[7 + 3 * c] & 15;
        comp[c].tq = seg[8 + 3 * c];
        if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) ERR(-1);
        if (comp[c].h > hmax) hmax = comp[c].h;
        if (comp[c].v > vmax) vmax = comp[c].v;
      }
      if (nc == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }  // a single component is never subsampled
      const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
      for (int c = 0; c < nc; ++c) {
        if ((comp[c].h != hmax && comp[c].h * 2 != hmax) || (comp[c].v != vmax && comp[c].v * 2 != vmax)) ERR(-2);
        comp[c].w = (W * comp[c].h + hmax - 1) / hmax;
        comp[c].h_px = (H * comp[c].v + vmax - 1) / vmax;
        comp[c].bw = mcux * comp[c].h;
        comp[c].bh = mcuy * comp[c].v;
        comp[c].coef = (int16_t *)calloc((size_t)comp[c].bw * comp[c].bh * 64, sizeof(int16_t));
        comp[c].plane = (uint8_t *)malloc((size_t)comp[c].bw * comp[c].bh * 64);
        if (!comp[c].coef || !comp[c].plane) ERR(-1);
      }
      have_sof = 1;
    } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
      ERR(-2);
    } else if (m == 0xDD) {
      if (sl < 2) ERR(-1);
      restart = (seg[0] << 8) | seg[1];
    } else if (m == 0xEE) {
      if (sl >= 12 && memcmp(seg, "Adobe", 5) == 0) adobe_transform = seg[11];
    } else if (m == 0xDA) {
      if (!have_sof || sl < 1) ERR(-1);
      const int ns = seg[0];
      if (ns < 1 || ns > nc || sl < 1 + 2 * (size_t)ns + 3) ERR(-1);
      Comp *sc_comp[4];
      for (int k = 0; k < ns; ++k) {
        Comp *c = NULL;
        for (int j = 0; j < nc; ++j)
          if (comp[j].id == seg[1 + 2 * k]) c = &comp[j];
        if (!c) ERR(-1);
        c->td = seg[2 + 2 * k] >> 4;
        c->ta = seg[2 + 2 * k] & 15;
        if (c->td > 3 || c->ta > 3) ERR(-1);
        sc_comp[k] = c;
      }
      Scan sc;
      sc.ss = seg[1 + 2 * ns];
      sc.se = seg[2 + 2 * ns];
      sc.ah = seg[3 + 2 * ns] >> 4;
      sc.al = seg[3 + 2 * ns] & 15;
      sc.progressive = progressive;
      sc.eobrun = 0;
      if (!progressive) { sc.ss = 0; sc.se = 63; sc.ah = sc.al = 0; }
      if (sc.ss > sc.se || sc.se > 63 || (progressive && sc.ss > 0 && ns != 1)) ERR(-1);
      Bits b;
      b.p = data + i + L;
      b.end = data + n;
      b.bitbuf = 0;
      b.bitcnt = 0;
      b.marker = 0;
      for (int j = 0; j < nc; ++j) comp[j].dc_pred = 0;
      // block iteration: interleaved = MCU order over all scan components; single component = its own block grid
      const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
      int units_x, units_y;
      if (ns == 1) {
        units_x = (sc_comp[0]->w + 7) / 8;
        units_y = (sc_comp[0]->h_px + 7) / 8;
      } else {
        units_x = mcux;
        units_y = mcuy;
      }
      int until_restart = restart;
      for (int uy = 0; uy < units_y; ++uy)
        for (int ux = 0; ux < units_x; ++ux) {
          if (restart && until_restart == 0) {
            // byte-align, expect RSTn
            b.bitcnt = 0;
            b.bitbuf = 0;
            if (!b.marker) {  // consume up to the marker
              while (b.p + 1 < b.end && !(b.p[0] == 0xFF && b.p[1] >= 0xD0 && b.p[1] <= 0xD7)) ++b.p;
              if (b.p + 1 < b.end) b.p += 2;
            } else if (b.marker < 0xD0 || b.marker > 0xD7) {
              ERR(-1);
            }
            b.marker = 0;
            for (int j = 0; j < nc; ++j) comp[j].dc_pred = 0;
            sc.eobrun = 0;
            until_restart = restart;
          }
          for (int k = 0; k < ns; ++k) {
            Comp *c = sc_comp[k];
            const int nh = ns == 1 ? 1 : c->h, nv = ns == 1 ? 1 : c->v;
            for (int by = 0; by < nv; ++by)
              for (int bx = 0; bx < nh; ++bx) {
                const int X = ux * nh + bx, Y = uy * nv + by;
                int16_t *coef = c->coef + ((size_t)Y * c->bw + X) * 64;
                const HuffTable *dc = &hdc[c->td], *ac = &hac[c->ta];
                if ((sc.ss == 0 && sc.ah == 0 && !dc->present) || (sc.se > 0 && !ac->present)) ERR(-1);
                int e = progressive ? decode_block_progressive(&b, c, dc, ac, coef, &sc) : decode_block_baseline(&b, c, dc, ac, coef);
                if (e) ERR(-1);
              }
          }
          if (restart) --until_restart;
        }
      // continue the marker loop after the entropy-coded data
      if (b.marker) {
        i = (size_t)(b.p - data) - 2;  // b.p sits just past the marker byte
      } else {
        const uint8_t *q = b.p;
        while (q + 1 < b.end && !(q[0] == 0xFF && q[1] != 0 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
        i = (size_t)(q - data);
      }
      continue;
    }
    i += L;
  }
  if (!have_sof) ERR(-1);
  if (nc == 3 && adobe_transform == 0) ERR(-2);  // RGB-coded JPEG: not produced by the reference's data, not restated
  // dequantise + IDCT every block of every component
  for (int c = 0; c < nc; ++c) {
    if (!qt_present[comp[c].tq]) ERR(-1);
    const int stride = comp[c].bw * 8;
    for (int by = 0; by < comp[c].bh; ++by)
      for (int bx = 0; bx < comp[c].bw; ++bx)
        idct_block(comp[c].coef + ((size_t)by * comp[c].bw + bx) * 64, qt[comp[c].tq], comp[c].plane + (size_t)by * 8 * stride + bx * 8, stride);
  }
  out = (uint8_t *)malloc((size_t)W * H * nc);
  if (!out) ERR(-1);
  if (nc == 1) {
    for (int y = 0; y < H; ++y) memcpy(out + (size_t)y * W, comp[0].plane + (size_t)y * comp[0].bw * 8, (size_t)W);
  } else {
    for (int c = 0; c < 3; ++c) {
      line[c] = (uint8_t *)malloc((size_t)comp[c].bw * 8 * 2 + 16);
      if (!line[c]) ERR(-1);
    }
    for (int y = 0; y < H; ++y) {
      for (int c = 0; c < 3; ++c) upsample_row(&comp[c], hmax, vmax, W, H, y, line[c]);
      for (int x = 0; x < W; ++x) ycbcr_to_rgb(variant, line[0][x], line[1][x], line[2][x], out + ((size_t)y * W + x) * 3);
    }
  }
  if (coef_out) {
    size_t total = 0;
    for (int c = 0; c < nc; ++c) total += (size_t)comp[c].bw * comp[c].bh * 64;
    int16_t *all = (int16_t *)malloc(total * sizeof(int16_t));
    if (!all) ERR(-1);
    size_t o = 0;
    for (int c = 0; c < nc; ++c) {
      memcpy(all + o, comp[c].coef, (size_t)comp[c].bw * comp[c].bh * 64 * sizeof(int16_t));
      o += (size_t)comp[c].bw * comp[c].bh * 64;
    }
    *coef_out = all;
    *coef_count = total;
  }
  *pixels = out;
  out = NULL;
  *width = W;
  *height = H;
  *n_comp = nc;
done:
  for (int c = 0; c < 4; ++c) {
    free(comp[c].coef);
    free(comp[c].plane);
    free(line[c]);
  }
  free(out);
  return rc;
}

int orc_jpeg_decode(const uint8_t *data, size_t n, int variant, uint8_t **pixels, int *width, int *height, int *n_comp) {
  return orc_jpeg_decode_ex(data, n, variant, pixels, width, height, n_comp, NULL, NULL);
}

void orc_decode_free(void *p) { free(p); }

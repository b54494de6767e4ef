// In-library multi-device entry point: the batch contract of the reference's evaluation loop
// (text_detection/mod.rs:188-204: images [n][H][W] + adjust [n][2] -> one PolygonScores for the whole batch) over
// several GPUs of one box, for hosts that are ONE process (the Rust binary is): images are independent, so the index
// range is cut into contiguous shards, one host thread per device runs the whole pipeline on its shard through that
// device's own ctx / detector / recognition net, and the results are appended into ONE host CSR in image order.
// No data-path collective exists or is needed; nothing here touches NCCL.
#include "common.cuh"

#include <string>
#include <thread>

namespace ocrb {
void polygons_append(ocrb_polygons *, const ocrb_polygons *);
ocrb_polygons *polygons_new();
}  // namespace ocrb

struct ocrb_shards {
  std::vector<int> devices;
  std::vector<ocrb_ctx *> ctx;
  std::vector<ocrb_det *> det;
  std::vector<ocrb_rec *> rec;
};

using namespace ocrb;

// contiguous [first, first + count) of shard r out of g; the first n % g shards hold one item more
static void shard_range(int64_t n, int r, int g, int64_t *first, int64_t *count) {
  const int64_t base = n / g, extra = n % g;
  *first = r * base + (r < extra ? r : extra);
  *count = base + (r < extra ? 1 : 0);
}

static int check_devices(const int *devices, int n_devices) {
  OCRB_REQUIRE(devices && n_devices > 0, "need at least one device");
  for (int i = 0; i < n_devices; ++i)
    for (int j = 0; j < i; ++j) OCRB_REQUIRE(devices[i] != devices[j], "device %d listed twice", devices[i]);
  return OCRB_OK;
}

extern "C" {

int ocrb_shards_destroy(ocrb_shards *s) {
  if (!s) return OCRB_OK;
  for (auto *r : s->rec) ocrb_rec_destroy(r);
  for (auto *d : s->det) ocrb_det_destroy(d);
  for (auto *c : s->ctx) ocrb_ctx_destroy(c);
  delete s;
  return OCRB_OK;
}

int ocrb_shards_create(const int *devices, int n_devices, int n_det, const char *const *det_names, const float *const *det_data,
                       const int64_t *det_numel, int mode, int n_rec, const char *const *rec_names, const float *const *rec_data,
                       const int64_t *rec_numel, ocrb_shards **out) {
  OCRB_REQUIRE(out, "null argument");
  OCRB_TRY(check_devices(devices, n_devices));
  ocrb_shards *s = new ocrb_shards();
  for (int i = 0; i < n_devices; ++i) {
    ocrb_ctx *c = nullptr;
    ocrb_det *d = nullptr;
    ocrb_rec *r = nullptr;
    int rc = ocrb_ctx_create(devices[i], &c);
    if (rc == OCRB_OK) {
      s->devices.push_back(devices[i]);
      s->ctx.push_back(c);
      rc = ocrb_det_create(c, n_det, det_names, det_data, det_numel, mode, &d);
    }
    if (rc == OCRB_OK) {
      s->det.push_back(d);
      if (n_rec > 0) rc = ocrb_rec_create(c, n_rec, rec_names, rec_data, rec_numel, &r);
    }
    if (rc == OCRB_OK && r) s->rec.push_back(r);
    if (rc != OCRB_OK) {
      ocrb_shards_destroy(s);
      return rc;
    }
  }
  *out = s;
  return OCRB_OK;
}

int ocrb_shards_create_from_files(const int *devices, int n_devices, const char *det_path, const char *rec_path, int mode, ocrb_shards **out) {
  OCRB_REQUIRE(out && det_path, "null argument");
  OCRB_TRY(check_devices(devices, n_devices));
  ocrb_shards *s = new ocrb_shards();
  for (int i = 0; i < n_devices; ++i) {
    ocrb_ctx *c = nullptr;
    ocrb_det *d = nullptr;
    ocrb_rec *r = nullptr;
    int rc = ocrb_ctx_create(devices[i], &c);
    if (rc == OCRB_OK) {
      s->devices.push_back(devices[i]);
      s->ctx.push_back(c);
      rc = ocrb_det_create_from_file(c, det_path, mode, &d);
    }
    if (rc == OCRB_OK) {
      s->det.push_back(d);
      if (rec_path) rc = ocrb_rec_create_from_file(c, rec_path, &r);
    }
    if (rc == OCRB_OK && r) s->rec.push_back(r);
    if (rc != OCRB_OK) {
      ocrb_shards_destroy(s);
      return rc;
    }
  }
  *out = s;
  return OCRB_OK;
}

int ocrb_shards_count(const ocrb_shards *s) { return s ? (int)s->devices.size() : 0; }
int ocrb_shards_device(const ocrb_shards *s, int i) { return (s && i >= 0 && i < (int)s->devices.size()) ? s->devices[i] : -1; }
int64_t ocrb_shards_launch_count(const ocrb_shards *s) {
  int64_t n = 0;
  if (s)
    for (auto *c : s->ctx) n += ocrb_ctx_launch_count(c);
  return n;
}

int ocrb_shard_range(int64_t n_items, int shard, int n_shards, int64_t *first, int64_t *count) {
  OCRB_REQUIRE(first && count && n_shards > 0 && shard >= 0 && shard < n_shards && n_items >= 0, "bad argument");
  shard_range(n_items, shard, n_shards, first, count);
  return OCRB_OK;
}

static int run_sharded(ocrb_shards *s, const uint8_t *images, const double *adjust, int B, int H, int W,
                       const ocrb_postproc_params *params, const uint8_t *glyphs, int n_glyphs, int32_t *glyph_argmax, int crop_k,
                       ocrb_polygons **out) {
  OCRB_REQUIRE(s && images && adjust && out, "null argument");
  OCRB_REQUIRE(crop_k == 0 || !s->rec.empty(), "glyph crops asked for without a recognition net");
  OCRB_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
  OCRB_REQUIRE(n_glyphs == 0 || (glyphs && !s->rec.empty()), "glyphs given without a recognition net");
  OCRB_REQUIRE(!is_device_ptr(images) && !is_device_ptr(glyphs) && !is_device_ptr(glyph_argmax),
               "the sharded entry point takes HOST buffers (pinned memory for full copy / compute overlap): one device pointer cannot feed several devices");
  const int G = (int)s->devices.size();
  const int64_t HW = (int64_t)H * W;
  struct Part {
    int rc = OCRB_OK;
    std::string err;
    ocrb_polygons *res = nullptr;
  };
  std::vector<Part> parts(G);
  std::vector<std::thread> threads;
  for (int r = 0; r < G; ++r) {
    int64_t first, count, gfirst, gcount;
    shard_range(B, r, G, &first, &count);
    shard_range(n_glyphs, r, G, &gfirst, &gcount);
    if (count == 0 && gcount == 0) continue;
    threads.emplace_back([=, &parts]() {
      Part &p = parts[r];
      if (count > 0 && crop_k > 0) {
        p.rc = ocrb_detect_and_read(s->det[r], s->rec[r], images + first * HW, adjust + first * 2, (int)count, H, W, params, crop_k, &p.res);
      } else if (count > 0) {
        p.rc = ocrb_detect_and_recognize(s->det[r], gcount > 0 ? s->rec[r] : nullptr, images + first * HW, adjust + first * 2, (int)count, H, W,
                                         params, gcount > 0 ? glyphs + gfirst * 784 : nullptr, (int)gcount,
                                         glyph_argmax && gcount > 0 ? glyph_argmax + gfirst : nullptr, &p.res);
      } else {  // more devices than images: this shard only classifies glyphs
        p.rc = ocrb_rec_forward_u8(s->rec[r], glyphs + gfirst * 784, (int)gcount, nullptr, glyph_argmax ? glyph_argmax + gfirst : nullptr, nullptr);
      }
      if (p.rc != OCRB_OK) p.err = ocrb_last_error();  // the message is thread-local: carry it to the caller's thread
    });
  }
  for (auto &t : threads) t.join();
  int rc = OCRB_OK;
  for (int r = 0; r < G && rc == OCRB_OK; ++r)
    if (parts[r].rc != OCRB_OK) {
      rc = parts[r].rc;
      set_error("shard %d (device %d): %s", r, s->devices[r], parts[r].err.c_str());
    }
  ocrb_polygons *res = nullptr;
  if (rc == OCRB_OK) {
    res = polygons_new();
    for (int r = 0; r < G; ++r)
      if (parts[r].res) polygons_append(res, parts[r].res);
  }
  for (auto &p : parts)
    if (p.res) ocrb_polygons_free(p.res);
  if (rc != OCRB_OK) return rc;
  *out = res;
  return OCRB_OK;
}

int ocrb_detect_and_recognize_sharded(ocrb_shards *s, const uint8_t *images, const double *adjust, int B, int H, int W,
                                      const ocrb_postproc_params *params, const uint8_t *glyphs, int n_glyphs,
                                      int32_t *glyph_argmax, ocrb_polygons **out) {
  return run_sharded(s, images, adjust, B, H, W, params, glyphs, n_glyphs, glyph_argmax, 0, out);
}

int ocrb_detect_and_read_sharded(ocrb_shards *s, const uint8_t *images, const double *adjust, int B, int H, int W,
                                 const ocrb_postproc_params *params, int glyphs_per_polygon, ocrb_polygons **out) {
  OCRB_REQUIRE(glyphs_per_polygon > 0, "glyphs_per_polygon must be positive");
  return run_sharded(s, images, adjust, B, H, W, params, nullptr, 0, nullptr, glyphs_per_polygon, out);
}

// page-locked host memory for the image / glyph buffers (full-speed, asynchronous H2D copies)
int ocrb_host_alloc(size_t bytes, void **out) {
  OCRB_REQUIRE(out && bytes > 0, "bad argument");
  OCRB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return OCRB_OK;
}
int ocrb_host_free(void *p) {
  if (p) OCRB_CUDA(cudaFreeHost(p));
  return OCRB_OK;
}

}  // extern "C"

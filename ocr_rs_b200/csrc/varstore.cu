// Native reader of the reference's model files: `vs.load(file)` (text_detection/mod.rs:40-44,
// char_recognition/mod.rs:43-45) reads what `VarStore::save` / `utils::save_vs` (utils.rs:55-63)
// wrote through tch 0.3.0 -> at_save_multi -> torch::serialize::OutputArchive: a ZIP archive
// (STORED entries) holding `<root>/data.pkl` — a protocol-2 pickle of one module object whose
// state is {variable name: _rebuild_tensor_v2(storage, offset, sizes, strides, ...)} — and the
// raw little-endian storages `<root>/data/<key>`.  This file parses exactly that (a ZIP
// central-directory walk + the dozen pickle opcodes the archive uses); no libtorch, no Python.
// Host-only code: it feeds ocrb_det_create / ocrb_rec_create.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace ocrb {
namespace {

struct ZipEntry {
  std::string name;
  uint64_t data_off = 0, size = 0;
  uint16_t method = 0;
};

uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

int zip_index(const std::vector<uint8_t> &f, std::vector<ZipEntry> &out) {
  const size_t n = f.size();
  if (n < 22) { set_error("model file too small to be a ZIP archive"); return OCRB_ERR_INVALID; }
  // end-of-central-directory record: scan backwards for PK\5\6
  size_t eocd = (size_t)-1;
  for (size_t i = n - 22;; --i) {
    if (rd32(&f[i]) == 0x06054b50u) { eocd = i; break; }
    if (i == 0 || n - i > 22 + 65535) break;
  }
  if (eocd == (size_t)-1) { set_error("model file is not a ZIP archive (no end-of-central-directory record)"); return OCRB_ERR_INVALID; }
  uint64_t count = rd16(&f[eocd + 10]), cd_size = rd32(&f[eocd + 12]), cd_off = rd32(&f[eocd + 16]);
  if (cd_off == 0xffffffffu || count == 0xffff) {  // ZIP64: locator just before the EOCD
    if (eocd < 20 || rd32(&f[eocd - 20]) != 0x07064b50u) { set_error("ZIP64 locator missing"); return OCRB_ERR_INVALID; }
    const uint64_t e64 = rd64(&f[eocd - 20 + 8]);
    if (e64 > n || n - e64 < 56 || rd32(&f[e64]) != 0x06064b50u) { set_error("bad ZIP64 end-of-central-directory record"); return OCRB_ERR_INVALID; }
    count = rd64(&f[e64 + 32]);
    cd_size = rd64(&f[e64 + 40]);
    cd_off = rd64(&f[e64 + 48]);
  }
  if (cd_off > n || cd_size > n - cd_off) { set_error("ZIP central directory out of range"); return OCRB_ERR_INVALID; }
  uint64_t p = cd_off;
  for (uint64_t i = 0; i < count; ++i) {
    if (p > n || n - p < 46 || rd32(&f[p]) != 0x02014b50u) { set_error("bad ZIP central-directory entry %llu", (unsigned long long)i); return OCRB_ERR_INVALID; }
    ZipEntry e;
    e.method = rd16(&f[p + 10]);
    uint64_t csize = rd32(&f[p + 20]), usize = rd32(&f[p + 24]), lho = rd32(&f[p + 42]);
    const uint16_t nl = rd16(&f[p + 28]), xl = rd16(&f[p + 30]), cl = rd16(&f[p + 32]);
    if (p + 46 + nl + xl + cl > n) { set_error("truncated ZIP central directory"); return OCRB_ERR_INVALID; }
    e.name.assign(reinterpret_cast<const char *>(&f[p + 46]), nl);
    // ZIP64 extra field (id 1): sizes / offset that overflowed 32 bits, in this order
    const uint8_t *x = &f[p + 46 + nl], *xe = x + xl;
    while (x + 4 <= xe) {
      const uint16_t id = rd16(x), len = rd16(x + 2);
      if (id == 1) {
        const uint8_t *q = x + 4;
        if (usize == 0xffffffffu && q + 8 <= xe) { usize = rd64(q); q += 8; }
        if (csize == 0xffffffffu && q + 8 <= xe) { csize = rd64(q); q += 8; }
        if (lho == 0xffffffffu && q + 8 <= xe) { lho = rd64(q); q += 8; }
      }
      x += 4 + len;
    }
    if (lho > n || n - lho < 30 || rd32(&f[lho]) != 0x04034b50u) { set_error("bad ZIP local header for %s", e.name.c_str()); return OCRB_ERR_INVALID; }
    e.data_off = lho + 30 + rd16(&f[lho + 26]) + rd16(&f[lho + 28]);
    e.size = usize;
    if (e.method == 0 && (e.data_off > n || e.size > n - e.data_off)) { set_error("ZIP entry %s out of range", e.name.c_str()); return OCRB_ERR_INVALID; }
    (void)csize;
    out.push_back(e);
    p += 46 + nl + xl + cl;
  }
  return OCRB_OK;
}

// ---- the pickle subset ------------------------------------------------------------------
struct Val;
typedef std::shared_ptr<Val> VP;
struct Val {
  enum Kind { NONE, INT, BOOL, FLOAT, STR, TUPLE, LIST, DICT, GLOBAL, STORAGE, TENSOR, OBJECT, MARK } kind = NONE;
  int64_t i = 0;
  double f = 0;
  std::string s;                      // STR; GLOBAL = "module name"; STORAGE = storage type
  std::vector<VP> items;              // TUPLE / LIST; DICT = k0, v0, k1, v1, ...
  // STORAGE: s = type, key, numel ; TENSOR: storage + offset, sizes, strides
  std::string key;
  int64_t numel = 0, offset = 0;
  std::vector<int64_t> sizes, strides;
  VP storage, state;                  // TENSOR.storage ; OBJECT.state
};
VP mk(Val::Kind k) { auto v = std::make_shared<Val>(); v->kind = k; return v; }

struct Unpickler {
  const uint8_t *p, *e;
  std::vector<VP> stack;
  std::map<uint32_t, VP> memo;
  bool need(size_t n) const { return (size_t)(e - p) >= n; }
  int fail(const char *what) { set_error("model file: unsupported or corrupt pickle (%s)", what); return OCRB_ERR_INVALID; }
  int pop_mark(std::vector<VP> &out) {
    size_t k = stack.size();
    while (k > 0 && stack[k - 1]->kind != Val::MARK) --k;
    if (k == 0) return fail("no MARK");
    out.assign(stack.begin() + k, stack.end());
    stack.resize(k - 1);
    return OCRB_OK;
  }
  static bool ints_of(const VP &t, std::vector<int64_t> &out) {
    if (!t || (t->kind != Val::TUPLE && t->kind != Val::LIST)) return false;
    for (auto &x : t->items) { if (x->kind != Val::INT) return false; out.push_back(x->i); }
    return true;
  }
  int reduce(const VP &fn, const VP &args, VP &out) {
    if (fn->kind != Val::GLOBAL || args->kind != Val::TUPLE) return fail("REDUCE of a non-global");
    if (fn->s == "torch._utils _rebuild_tensor_v2" || fn->s == "torch._utils _rebuild_tensor") {
      if (args->items.size() < 4 || args->items[0]->kind != Val::STORAGE || args->items[1]->kind != Val::INT) return fail("_rebuild_tensor arguments");
      out = mk(Val::TENSOR);
      out->storage = args->items[0];
      out->offset = args->items[1]->i;
      if (!ints_of(args->items[2], out->sizes) || !ints_of(args->items[3], out->strides)) return fail("tensor sizes/strides");
      return OCRB_OK;
    }
    if (fn->s == "collections OrderedDict") { out = mk(Val::DICT); return OCRB_OK; }
    if (fn->s == "torch._utils _rebuild_parameter") {  // (tensor, requires_grad, backward_hooks)
      if (args->items.empty() || args->items[0]->kind != Val::TENSOR) return fail("_rebuild_parameter arguments");
      out = args->items[0];
      return OCRB_OK;
    }
    return fail(fn->s.c_str());
  }
  int run(VP &result) {
    while (p < e) {
      const uint8_t op = *p++;
      switch (op) {
        case 0x80: if (!need(1)) return fail("PROTO"); ++p; break;                                   // PROTO
        case '.': if (stack.empty()) return fail("STOP on empty stack"); result = stack.back(); return OCRB_OK;
        case '(': stack.push_back(mk(Val::MARK)); break;
        case 'N': stack.push_back(mk(Val::NONE)); break;
        case 0x88: case 0x89: { auto v = mk(Val::BOOL); v->i = op == 0x88; stack.push_back(v); break; }
        case 'K': { if (!need(1)) return fail("BININT1"); auto v = mk(Val::INT); v->i = *p++; stack.push_back(v); break; }
        case 'M': { if (!need(2)) return fail("BININT2"); auto v = mk(Val::INT); v->i = rd16(p); p += 2; stack.push_back(v); break; }
        case 'J': { if (!need(4)) return fail("BININT"); auto v = mk(Val::INT); v->i = (int32_t)rd32(p); p += 4; stack.push_back(v); break; }
        case 0x8a: {                                                                                  // LONG1
          if (!need(1)) return fail("LONG1");
          const int n = *p++;
          if (n > 8 || !need((size_t)n)) return fail("LONG1 size");
          uint64_t u = 0;
          for (int k = 0; k < n; ++k) u |= (uint64_t)p[k] << (8 * k);
          if (n > 0 && n < 8 && (p[n - 1] & 0x80)) u |= ~(uint64_t)0 << (8 * n);
          p += n;
          auto v = mk(Val::INT); v->i = (int64_t)u; stack.push_back(v); break;
        }
        case 'G': {                                                                                   // BINFLOAT (big endian)
          if (!need(8)) return fail("BINFLOAT");
          uint64_t u = 0;
          for (int k = 0; k < 8; ++k) u = (u << 8) | p[k];
          p += 8;
          auto v = mk(Val::FLOAT); memcpy(&v->f, &u, 8); stack.push_back(v); break;
        }
        case 'X': {                                                                                   // BINUNICODE
          if (!need(4)) return fail("BINUNICODE");
          const uint32_t n = rd32(p); p += 4;
          if (!need(n)) return fail("BINUNICODE length");
          auto v = mk(Val::STR); v->s.assign(reinterpret_cast<const char *>(p), n); p += n; stack.push_back(v); break;
        }
        case 0x8c: {                                                                                  // SHORT_BINUNICODE
          if (!need(1)) return fail("SHORT_BINUNICODE");
          const uint32_t n = *p++;
          if (!need(n)) return fail("SHORT_BINUNICODE length");
          auto v = mk(Val::STR); v->s.assign(reinterpret_cast<const char *>(p), n); p += n; stack.push_back(v); break;
        }
        case 'c': {                                                                                   // GLOBAL "module\nname\n"
          const uint8_t *a = p;
          while (p < e && *p != '\n') ++p;
          if (p >= e) return fail("GLOBAL");
          std::string mod(reinterpret_cast<const char *>(a), p - a);
          a = ++p;
          while (p < e && *p != '\n') ++p;
          if (p >= e) return fail("GLOBAL");
          std::string nm(reinterpret_cast<const char *>(a), p - a);
          ++p;
          auto v = mk(Val::GLOBAL); v->s = mod + " " + nm; stack.push_back(v); break;
        }
        case 'q': { if (!need(1) || stack.empty()) return fail("BINPUT"); memo[*p++] = stack.back(); break; }
        case 'r': { if (!need(4) || stack.empty()) return fail("LONG_BINPUT"); memo[rd32(p)] = stack.back(); p += 4; break; }
        case 0x94: { if (stack.empty()) return fail("MEMOIZE"); const uint32_t k = (uint32_t)memo.size(); memo[k] = stack.back(); break; }
        case 'h': { if (!need(1)) return fail("BINGET"); auto it = memo.find(*p++); if (it == memo.end()) return fail("BINGET key"); stack.push_back(it->second); break; }
        case 'j': { if (!need(4)) return fail("LONG_BINGET"); auto it = memo.find(rd32(p)); p += 4; if (it == memo.end()) return fail("LONG_BINGET key"); stack.push_back(it->second); break; }
        case ')': stack.push_back(mk(Val::TUPLE)); break;
        case ']': stack.push_back(mk(Val::LIST)); break;
        case '}': stack.push_back(mk(Val::DICT)); break;
        case 't': { auto v = mk(Val::TUPLE); OCRB_TRY(pop_mark(v->items)); stack.push_back(v); break; }
        case 0x85: case 0x86: case 0x87: {                                                            // TUPLE1..3
          const size_t n = op - 0x84;
          if (stack.size() < n) return fail("TUPLEn");
          auto v = mk(Val::TUPLE); v->items.assign(stack.end() - n, stack.end()); stack.resize(stack.size() - n); stack.push_back(v); break;
        }
        case 'a': { if (stack.size() < 2) return fail("APPEND"); VP x = stack.back(); stack.pop_back(); stack.back()->items.push_back(x); break; }
        case 'e': { std::vector<VP> xs; OCRB_TRY(pop_mark(xs)); if (stack.empty()) return fail("APPENDS"); for (auto &x : xs) stack.back()->items.push_back(x); break; }
        case 's': { if (stack.size() < 3) return fail("SETITEM"); VP v = stack.back(); stack.pop_back(); VP k = stack.back(); stack.pop_back(); stack.back()->items.push_back(k); stack.back()->items.push_back(v); break; }
        case 'u': { std::vector<VP> xs; OCRB_TRY(pop_mark(xs)); if (stack.empty() || xs.size() % 2) return fail("SETITEMS"); for (auto &x : xs) stack.back()->items.push_back(x); break; }
        case 'Q': {                                                                                   // BINPERSID: ('storage', type, key, location, numel)
          if (stack.empty()) return fail("BINPERSID");
          VP t = stack.back(); stack.pop_back();
          if (t->kind != Val::TUPLE || t->items.size() < 5 || t->items[0]->kind != Val::STR || t->items[0]->s != "storage" ||
              t->items[1]->kind != Val::GLOBAL || t->items[2]->kind != Val::STR || t->items[4]->kind != Val::INT)
            return fail("persistent id is not a storage");
          auto v = mk(Val::STORAGE); v->s = t->items[1]->s; v->key = t->items[2]->s; v->numel = t->items[4]->i; stack.push_back(v); break;
        }
        case 'R': { if (stack.size() < 2) return fail("REDUCE"); VP a = stack.back(); stack.pop_back(); VP f = stack.back(); stack.pop_back(); VP o; OCRB_TRY(reduce(f, a, o)); stack.push_back(o); break; }
        case 0x81: { if (stack.size() < 2) return fail("NEWOBJ"); stack.pop_back(); stack.pop_back(); stack.push_back(mk(Val::OBJECT)); break; }
        case 'b': { if (stack.size() < 2) return fail("BUILD"); VP st = stack.back(); stack.pop_back(); if (stack.back()->kind == Val::OBJECT) stack.back()->state = st; break; }
        default: { char b[32]; snprintf(b, sizeof(b), "opcode 0x%02x", op); return fail(b); }
      }
    }
    return fail("no STOP");
  }
};

int elem_size(const std::string &storage_type, int *kind) {
  // kind: 0 f32, 1 f64, 2 f16, 3 bf16, 4 i64, 5 i32, 6 u8
  static const struct { const char *n; int sz, k; } T[] = {{"torch FloatStorage", 4, 0}, {"torch DoubleStorage", 8, 1}, {"torch HalfStorage", 2, 2},
                                                          {"torch BFloat16Storage", 2, 3}, {"torch LongStorage", 8, 4}, {"torch IntStorage", 4, 5},
                                                          {"torch ByteStorage", 1, 6}};
  for (auto &t : T)
    if (storage_type == t.n) { *kind = t.k; return t.sz; }
  return 0;
}

float half_to_float(uint16_t h) {
  const uint32_t s = (h >> 15) & 1, ex = (h >> 10) & 31, m = h & 1023;
  uint32_t u;
  if (ex == 0) {
    if (m == 0) u = s << 31;
    else { int e2 = -1; uint32_t mm = m; do { ++e2; mm <<= 1; } while (!(mm & 1024)); u = (s << 31) | ((uint32_t)(127 - 15 - e2) << 23) | ((mm & 1023) << 13); }
  } else if (ex == 31) u = (s << 31) | 0x7f800000u | (m << 13);
  else u = (s << 31) | ((ex + 112) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}

}  // namespace
}  // namespace ocrb

struct ocrb_varstore {
  std::vector<std::string> names;
  std::vector<std::vector<float>> data;
  std::vector<std::vector<int64_t>> shapes;
};

using namespace ocrb;

static int varstore_open_impl(const char *path, ocrb_varstore **out) {
  OCRB_REQUIRE(path && out, "null argument");
  FILE *fp = fopen(path, "rb");
  if (!fp) { set_error("cannot open model file %s", path); return OCRB_ERR_INVALID; }
  std::vector<uint8_t> f;
  {
    fseek(fp, 0, SEEK_END);
    const long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (n < 0) { fclose(fp); set_error("cannot size %s", path); return OCRB_ERR_INVALID; }
    f.resize((size_t)n);
    const size_t got = fread(f.data(), 1, f.size(), fp);
    fclose(fp);
    if (got != f.size()) { set_error("short read of %s", path); return OCRB_ERR_INVALID; }
  }
  std::vector<ZipEntry> entries;
  OCRB_TRY(zip_index(f, entries));
  const ZipEntry *pkl = nullptr;
  std::string root;
  for (auto &e : entries) {
    const size_t k = e.name.rfind("/data.pkl");
    if (e.name == "data.pkl" || (k != std::string::npos && k + 9 == e.name.size() && e.name.find('/') == k)) {
      pkl = &e;
      root = e.name.substr(0, e.name.size() - 8);  // "<root>/" or ""
    }
  }
  OCRB_REQUIRE(pkl, "%s has no data.pkl: not a libtorch archive", path);
  OCRB_REQUIRE(pkl->method == 0, "compressed ZIP entries are not supported (libtorch writes STORED entries)");
  Unpickler up{f.data() + pkl->data_off, f.data() + pkl->data_off + pkl->size, {}, {}};
  VP top;
  OCRB_TRY(up.run(top));
  // module object -> state dict; a bare dict (torch.save of a state_dict) is accepted too
  VP dict = top->kind == Val::OBJECT ? top->state : top;
  OCRB_REQUIRE(dict && dict->kind == Val::DICT, "%s: the archive's top-level object has no variable dictionary", path);
  std::unique_ptr<ocrb_varstore> vs(new ocrb_varstore());
  for (size_t i = 0; i + 1 < dict->items.size(); i += 2) {
    const VP &k = dict->items[i], &v = dict->items[i + 1];
    if (k->kind != Val::STR || v->kind != Val::TENSOR) continue;  // e.g. `training` flags
    int kind = 0;
    const int esz = elem_size(v->storage->s, &kind);
    OCRB_REQUIRE(esz > 0, "variable %s has unsupported storage type %s", k->s.c_str(), v->storage->s.c_str());
    const ZipEntry *blob = nullptr;
    const std::string want = root + "data/" + v->storage->key;
    for (auto &e : entries)
      if (e.name == want) blob = &e;
    OCRB_REQUIRE(blob && blob->method == 0, "storage %s of variable %s is missing from the archive", want.c_str(), k->s.c_str());
    // sizes come from the pickle: bound the element count before allocating.  A tensor may be a broadcast view
    // (stride 0) of a smaller storage, so the bound is generous but finite: 2^31 elements.
    int64_t numel = 1;
    bool sane = v->sizes.size() == v->strides.size() && v->sizes.size() <= 8;
    for (int64_t d : v->sizes) {
      if (d < 0 || (d > 0 && numel > ((int64_t)1 << 31) / d)) { sane = false; break; }
      numel *= d;
    }
    OCRB_REQUIRE(sane, "variable %s has inconsistent or oversized sizes/strides", k->s.c_str());
    const uint64_t storage_elems = blob->size / (uint64_t)esz;
    std::vector<float> vals((size_t)numel);
    const uint8_t *base = f.data() + blob->data_off;
    const int nd = (int)v->sizes.size();
    std::vector<int64_t> idx(nd, 0);
    for (int64_t n = 0; n < numel; ++n) {  // general strided gather (tch writes contiguous tensors; views are legal)
      int64_t off = v->offset;
      for (int d = 0; d < nd; ++d) off += idx[d] * v->strides[d];
      OCRB_REQUIRE(off >= 0 && (uint64_t)off < storage_elems, "variable %s reads outside its storage", k->s.c_str());
      const uint8_t *q = base + off * esz;
      float x;
      switch (kind) {
        case 0: memcpy(&x, q, 4); break;
        case 1: { double dd; memcpy(&dd, q, 8); x = (float)dd; break; }
        case 2: x = half_to_float(rd16(q)); break;
        case 3: { uint32_t u = (uint32_t)rd16(q) << 16; memcpy(&x, &u, 4); break; }
        case 4: { int64_t ll; memcpy(&ll, q, 8); x = (float)ll; break; }
        case 5: { int32_t ii; memcpy(&ii, q, 4); x = (float)ii; break; }
        default: x = (float)*q; break;
      }
      vals[(size_t)n] = x;
      for (int d = nd - 1; d >= 0; --d) { if (++idx[d] < v->sizes[d]) break; idx[d] = 0; }
    }
    vs->names.push_back(k->s);
    vs->data.push_back(std::move(vals));
    vs->shapes.push_back(v->sizes);
  }
  OCRB_REQUIRE(!vs->names.empty(), "%s holds no tensors", path);
  *out = vs.release();
  return OCRB_OK;
}

extern "C" {

// nothing throws across the ABI: a corrupt archive can still make a container throw (bad_alloc, length_error)
int ocrb_varstore_open(const char *path, ocrb_varstore **out) {
  try {
    return varstore_open_impl(path, out);
  } catch (const std::exception &e) {
    set_error("model file %s: %s", path ? path : "(null)", e.what());
    return OCRB_ERR_INVALID;
  } catch (...) {
    set_error("model file %s: unknown failure", path ? path : "(null)");
    return OCRB_ERR_INVALID;
  }
}

int ocrb_varstore_count(const ocrb_varstore *vs) { return vs ? (int)vs->names.size() : 0; }
const char *ocrb_varstore_name(const ocrb_varstore *vs, int i) { return (vs && i >= 0 && i < (int)vs->names.size()) ? vs->names[i].c_str() : nullptr; }
int ocrb_varstore_tensor(const ocrb_varstore *vs, int i, const float **data, int64_t *numel, const int64_t **shape, int *ndim) {
  OCRB_REQUIRE(vs && i >= 0 && i < (int)vs->names.size(), "variable index out of range");
  if (data) *data = vs->data[i].data();
  if (numel) *numel = (int64_t)vs->data[i].size();
  if (shape) *shape = vs->shapes[i].data();
  if (ndim) *ndim = (int)vs->shapes[i].size();
  return OCRB_OK;
}
void ocrb_varstore_close(ocrb_varstore *vs) { delete vs; }

static void varstore_args(const ocrb_varstore *vs, std::vector<const char *> &names, std::vector<const float *> &data, std::vector<int64_t> &numel) {
  for (size_t i = 0; i < vs->names.size(); ++i) {
    names.push_back(vs->names[i].c_str());
    data.push_back(vs->data[i].data());
    numel.push_back((int64_t)vs->data[i].size());
  }
}

int ocrb_det_create_from_file(ocrb_ctx *ctx, const char *path, int mode, ocrb_det **out) {
  ocrb_varstore *vs = nullptr;
  OCRB_TRY(ocrb_varstore_open(path, &vs));
  std::vector<const char *> names;
  std::vector<const float *> data;
  std::vector<int64_t> numel;
  varstore_args(vs, names, data, numel);
  const int rc = ocrb_det_create(ctx, (int)names.size(), names.data(), data.data(), numel.data(), mode, out);
  ocrb_varstore_close(vs);
  return rc;
}

int ocrb_rec_create_from_file(ocrb_ctx *ctx, const char *path, ocrb_rec **out) {
  ocrb_varstore *vs = nullptr;
  OCRB_TRY(ocrb_varstore_open(path, &vs));
  std::vector<const char *> names;
  std::vector<const float *> data;
  std::vector<int64_t> numel;
  varstore_args(vs, names, data, numel);
  const int rc = ocrb_rec_create(ctx, (int)names.size(), names.data(), data.data(), numel.data(), out);
  ocrb_varstore_close(vs);
  return rc;
}

}  // extern "C"

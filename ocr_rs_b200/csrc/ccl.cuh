// Shared view of the connected-component labels (ccl.cu) for their consumers (contours.cu).
//
// labels[i] is a union-find parent (image-local linear index); roots are final after the seam pass.  Tiles of
// CCL_TW x CCL_TH pixels WITHOUT a foreground pixel are one background rectangle whose pixels all point at the tile
// origin: the local pass writes ONLY the origin's label there (4 B per tile instead of 4 KB — most of a document page),
// so every read of an arbitrary pixel's label goes through ccl_parent(), which answers "the origin" for the other
// pixels of such a tile without touching memory.
#pragma once
#include <cstdint>

namespace ocrb {

constexpr int CCL_TW = 32;  // tile width  (= warp size: one warp per tile row)
constexpr int CCL_TH = 32;  // tile height (64 measured: local pass 4.1 -> 4.3 ms, seam 1.26 -> 1.15 ms per 1024 images: no gain)

struct CclTiles {
  const uint8_t *empty;  // [B][tiles_y][tiles_x]: 1 = the tile holds no foreground pixel
  int tiles_x, tiles_per_img;
};

#ifdef __CUDACC__
// parent of pixel i of image b (L = that image's label plane)
__device__ __forceinline__ int ccl_parent(const int *__restrict__ L, const CclTiles &t, int b, int i, int W) {
  const int x = i % W, y = i / W;
  const int tx = x / CCL_TW, ty = y / CCL_TH;
  if (t.empty[(int64_t)b * t.tiles_per_img + ty * t.tiles_x + tx]) {
    const int origin = ty * CCL_TH * W + tx * CCL_TW;
    if (i != origin) return origin;
  }
  return L[i];
}
// root of an arbitrary pixel: after the first hop the walk only visits roots, whose labels are always stored
__device__ __forceinline__ int ccl_find_px(const int *__restrict__ L, const CclTiles &t, int b, int i, int W) {
  int a = i, p = ccl_parent(L, t, b, i, W);
  while (p != a) {
    a = p;
    p = L[a];
  }
  return a;
}
#endif

}  // namespace ocrb

// Host-side helpers shared by the TMA-fed kernels: driver entry point for tensor-map encoding, tile-box geometry and a
// per-thread tensor-map cache that hands out copies.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <map>

namespace dirb200 {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// One M tile = 128 consecutive output pixels = a {wbox, hbox, nbox} brick of the (Wo, Ho, B) output grid.
struct Boxes {
  int wbox, hbox, nbox;
};
inline Boxes pick_boxes(int Ho, int Wo, int bm = 128) {
  Boxes b;
  b.wbox = Wo < bm ? Wo : bm;
  b.hbox = Ho < bm / b.wbox ? Ho : bm / b.wbox;
  b.nbox = bm / (b.wbox * b.hbox);
  return b;
}

// Lookups hand out COPIES: an insertion may evict (clear) the cache, so a pointer into it could dangle while the same
// launch is still collecting its other maps.
template <typename Key>
struct MapCache {
  std::map<Key, CUtensorMap> m;
  bool find(const Key& k, CUtensorMap* out) const {
    auto it = m.find(k);
    if (it == m.end()) return false;
    *out = it->second;
    return true;
  }
  void put(const Key& k, const CUtensorMap& v) {
    if (m.size() > 8192) m.clear();
    m.emplace(k, v);
  }
};

}  // namespace tma
}  // namespace dirb200

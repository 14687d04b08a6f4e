// Host-side helpers shared by the tcgen05 launchers: the driver's tensor-map encoder, a cache of encoded
// CUtensorMaps keyed by (pointer, geometry) -- include/snnqp.h promises that the hot call neither allocates nor
// re-derives descriptors -- and a once-per-device cudaFuncSetAttribute.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace snnqp {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void *ptr;
  int32_t g[6];        // geometry + a tag that tells the encodings of one launcher apart
  int64_t s[2];        // outer strides
  bool operator==(const TmapKey &o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey &k) const {
    uint64_t h = 1469598103934665603ull;
    const unsigned char *b = reinterpret_cast<const unsigned char *>(&k);
    for (size_t i = 0; i < sizeof(TmapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return (size_t)h;
  }
};

// Returns the cached tensor map for `key`, encoding it with `make(CUtensorMap*) -> CUresult` on a miss; nullptr if
// the encoder failed.  Entries live until the cache (512 maps, 64 KB) wraps; the returned object is a per-thread
// copy, valid until this thread's next call with the same slot (two slots: x / w of one launch).
template <typename Make>
const CUtensorMap *tmap_cache_get(const TmapKey &key_in, Make make) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static thread_local CUtensorMap out[2];
  static thread_local int slot = 0;
  TmapKey key;
  memset(&key, 0, sizeof(key));            // padding bytes take part in hashing / comparison
  key.ptr = key_in.ptr;
  memcpy(key.g, key_in.g, sizeof(key.g));
  memcpy(key.s, key_in.s, sizeof(key.s));
  CUtensorMap *o = &out[slot];
  slot ^= 1;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap tm;
    if (make(&tm) != CUDA_SUCCESS) return nullptr;
    if (cache.size() >= 512) cache.clear();
    it = cache.emplace(key, tm).first;
  }
  *o = it->second;
  return o;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel, device) instead of on every launch
template <auto Kernel>
int ensure_smem_attr(int bytes) {
  static bool done[64] = {};
  int dev = 0;
  SNNQP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    SNNQP_CUDA(cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return SNNQP_OK;
}

}  // namespace snnqp

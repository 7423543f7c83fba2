// Stand-in for <cub/cub.cuh> in the CPU emulation build (tools/cuemu; tests only): the two device-wide
// primitives the library calls, with CUB's two-phase (size query, run) calling convention.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../../cuemu.h"

namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void *tmp, size_t &tmp_bytes, const K *kin, K *kout, const V *vin, V *vout, int n,
                               int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) { tmp_bytes = 256; return cudaSuccess; }
    const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8) ? ~K(0) : (K)(((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (kin[a] & mask) < (kin[b] & mask); });
    std::vector<K> ks((size_t)n);
    std::vector<V> vs((size_t)n);
    for (int i = 0; i < n; i++) { ks[i] = kin[idx[i]]; vs[i] = vin[idx[i]]; }
    for (int i = 0; i < n; i++) { kout[i] = ks[i]; vout[i] = vs[i]; }
    return cudaSuccess;
  }
};
struct DeviceScan {
  template <class T>
  static cudaError_t ExclusiveSum(void *tmp, size_t &tmp_bytes, const T *in, T *out, int n, cudaStream_t = nullptr) {
    if (!tmp) { tmp_bytes = 256; return cudaSuccess; }
    T run = 0;
    for (int i = 0; i < n; i++) { const T v = in[i]; out[i] = run; run += v; }
    return cudaSuccess;
  }
};
}  // namespace cub

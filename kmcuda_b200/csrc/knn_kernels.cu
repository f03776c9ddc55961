// knn_kernels.cu -- cluster-pruned exact k-NN (reference knn.cu:19-347), row-major samples.
//
// The search keeps the reference's decision arithmetic: true distances by Kahan-compensated
// round-down FMA sums (exact.cuh), cluster skip test `Cd[B][A] - d(q,A) - R[B] > kth`
// (knn.cu:218-225), insertion on `dist <= kth` into a binary max-heap (knn.cu:133-175), output in
// ascending distance order by popping (knn.cu:239-242).
#include "exact.cuh"
#include "kernels.h"

namespace kmb {

static inline unsigned cdivk(size_t a, size_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// distance accumulated over feature chunks: fresh Kahan sum per chunk, chunks added with a plain
// fp32 add, finalize at the end (knn.cu:36-47 with 16-feature chunks, knn.cu:79-100 with 24)
template <int METRIC, int CHUNK>
__device__ __forceinline__ float chunked_distance(const float* __restrict__ a,
                                                  const float* __restrict__ b, int D) {
  float acc = 0.f;
  for (int f0 = 0; f0 < D; f0 += CHUNK) {
    int fl = min(CHUNK, D - f0);
    Kahan k;
    if (METRIC == 1) for (int f = 0; f < fl; f++) k.mac(a[f0 + f], b[f0 + f]);
    else for (int f = 0; f < fl; f++) k.sqdiff(a[f0 + f], b[f0 + f]);
    acc += k.sum;
  }
  return finalize_distance<METRIC>(acc);
}

template <int METRIC>
__global__ void knn_radii_kernel(const float* __restrict__ X, const float* __restrict__ C, uint32_t n,
                                 int D, uint32_t K, const uint32_t* __restrict__ assign,
                                 uint32_t* __restrict__ radii_bits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a = assign[i];
  if (a >= K) return;
  float d = chunked_distance<METRIC, 16>(X + static_cast<size_t>(i) * D, C + static_cast<size_t>(a) * D, D);
  if (d == d) atomicMax(radii_bits + a, __float_as_uint(fmaxf(d, 0.f)));
}

__global__ void knn_radii_fix_kernel(const uint32_t* __restrict__ assign_counts_off, uint32_t K,
                                     float* __restrict__ radii) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= K) return;
  if (assign_counts_off[c + 1] == assign_counts_off[c]) radii[c] = __int_as_float(0x7fc00000);
}

cudaError_t launch_knn_radii(int metric, const float* X, const float* C, uint32_t n, int D, uint32_t K,
                             const uint32_t* assign, float* radii, cudaStream_t st) {
  cudaMemsetAsync(radii, 0, sizeof(float) * K, st);
  if (n == 0) return cudaGetLastError();
  if (metric == 1)
    knn_radii_kernel<1><<<cdivk(n, 256), 256, 0, st>>>(X, C, n, D, K, assign, reinterpret_cast<uint32_t*>(radii));
  else
    knn_radii_kernel<0><<<cdivk(n, 256), 256, 0, st>>>(X, C, n, D, K, assign, reinterpret_cast<uint32_t*>(radii));
  return cudaGetLastError();
}

template <int METRIC>
__global__ void knn_cdist_kernel(const float* __restrict__ C, uint32_t K, int D, float* __restrict__ cd) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(K) * K) return;
  uint32_t a = i / K, b = i - static_cast<size_t>(a) * K;
  // the reference fills the upper triangle and mirrors it (knn.cu:60-131): evaluate (min,max) order
  uint32_t lo = min(a, b), hi = max(a, b);
  cd[i] = chunked_distance<METRIC, 24>(C + static_cast<size_t>(lo) * D, C + static_cast<size_t>(hi) * D, D);
}

cudaError_t launch_knn_centroid_distances(int metric, const float* C, uint32_t K, int D, float* cd,
                                          cudaStream_t st) {
  size_t n = static_cast<size_t>(K) * K;
  if (metric == 1) knn_cdist_kernel<1><<<cdivk(n, 128), 128, 0, st>>>(C, K, D, cd);
  else knn_cdist_kernel<0><<<cdivk(n, 128), 128, 0, st>>>(C, K, D, cd);
  return cudaGetLastError();
}

// binary max-heap in a strided scratch column: element s of query q lives at hp[(2s)*stride],
// its index at hp[(2s+1)*stride]  (reference push_sample, knn.cu:133-175)
__device__ __forceinline__ void heap_push(int k, float dist, uint32_t index, float* hp, size_t stride) {
  int pos = 0;
  for (;;) {
    float left = 0.f, right = 0.f;
    bool left_le, right_le;
    if (2 * pos + 1 < k) {
      left = hp[(2 * (2 * pos + 1)) * stride];
      left_le = dist >= left;
    } else {
      left_le = true;
    }
    if (2 * pos + 2 < k) {
      right = hp[(2 * (2 * pos + 2)) * stride];
      right_le = dist >= right;
    } else {
      right_le = true;
    }
    if (left_le && right_le) {
      hp[(2 * pos) * stride] = dist;
      hp[(2 * pos + 1) * stride] = __uint_as_float(index);
      return;
    }
    bool go_right = (!left_le && !right_le) ? (left <= right) : left_le;
    int child = go_right ? 2 * pos + 2 : 2 * pos + 1;
    hp[(2 * pos) * stride] = hp[(2 * child) * stride];
    hp[(2 * pos + 1) * stride] = hp[(2 * child + 1) * stride];
    pos = child;
  }
}

template <int METRIC>
__global__ void __launch_bounds__(128)
knn_search_kernel(int k, const float* __restrict__ X, const float* __restrict__ C, uint32_t N, int D,
                  uint32_t K, uint32_t q_offset, uint32_t q_length, const uint32_t* __restrict__ assign,
                  const uint32_t* __restrict__ inv, const uint32_t* __restrict__ inv_off,
                  const float* __restrict__ cd, const float* __restrict__ radii,
                  float* __restrict__ heap_scratch, uint32_t* __restrict__ neighbors,
                  unsigned long long* __restrict__ d_pairs,
                  // list mode (rows != nullptr): the queries are rows[0 .. *d_nrows), heap column = list slot
                  const uint32_t* __restrict__ rows, const uint32_t* __restrict__ d_nrows) {
  const uint32_t count = rows ? min(*d_nrows, q_length) : q_length;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
  const uint32_t s = rows ? rows[q] : q_offset + q;
  const uint32_t oq = s - q_offset;          // output row
  const float* xs = X + static_cast<size_t>(s) * D;
  const size_t stride = q_length;
  float* hp = heap_scratch + q;
  for (int i = 0; i < k; i++) {
    hp[(2 * i) * stride] = FLT_MAX;
    hp[(2 * i + 1) * stride] = __uint_as_float(UINT32_MAX);
  }
  const uint32_t A = assign[s];
  unsigned long long pairs = 0;
  float kth = FLT_MAX;
  float dA = 0.f;
  if (A < K) {
    dA = distance_exact<METRIC>(xs, C + static_cast<size_t>(A) * D, D);
    uint32_t b = inv_off[A], e = inv_off[A + 1];
    pairs += e - b;
    for (uint32_t p = b; p < e; p++) {
      uint32_t o = inv[p];
      if (o == s) continue;
      float d = distance_exact<METRIC>(xs, X + static_cast<size_t>(o) * D, D);
      if (d <= kth) {
        heap_push(k, d, o, hp, stride);
        kth = hp[0];
      }
    }
  }
  for (uint32_t B = 0; B < K; B++) {
    if (B == A) continue;
    float cdist = A < K ? cd[static_cast<size_t>(B) * K + A] : 0.f;
    if (cdist != cdist) continue;
    if (A < K && cdist - dA - radii[B] > kth) continue;
    uint32_t b = inv_off[B], e = inv_off[B + 1];
    pairs += e - b;
    for (uint32_t p = b; p < e; p++) {
      uint32_t o = inv[p];
      float d = distance_exact<METRIC>(xs, X + static_cast<size_t>(o) * D, D);
      if (d <= kth) {
        heap_push(k, d, o, hp, stride);
        kth = hp[0];
      }
    }
  }
  for (int i = k - 1; i >= 0; i--) {
    neighbors[static_cast<size_t>(oq) * k + i] = __float_as_uint(hp[stride]);
    heap_push(k, -1.f, UINT32_MAX, hp, stride);
  }
  atomicAdd(d_pairs, pairs);
  }
}

cudaError_t launch_knn_search(int metric, int k, const float* X, const float* C, uint32_t N, int D,
                              uint32_t K, uint32_t q_offset, uint32_t q_length, const uint32_t* assign,
                              const uint32_t* inv, const uint32_t* inv_off, const float* cd,
                              const float* radii, float* heap_scratch, uint32_t* neighbors,
                              unsigned long long* d_pairs, const uint32_t* rows, const uint32_t* d_nrows,
                              cudaStream_t st) {
  if (q_length == 0) return cudaSuccess;
  const unsigned grid = rows ? 148u * 8u : cdivk(q_length, 128);
  if (metric == 1)
    knn_search_kernel<1><<<grid, 128, 0, st>>>(k, X, C, N, D, K, q_offset, q_length, assign, inv, inv_off, cd, radii,
                                               heap_scratch, neighbors, d_pairs, rows, d_nrows);
  else
    knn_search_kernel<0><<<grid, 128, 0, st>>>(k, X, C, N, D, K, q_offset, q_length, assign, inv, inv_off, cd, radii,
                                               heap_scratch, neighbors, d_pairs, rows, d_nrows);
  return cudaGetLastError();
}

// empty clusters have no radius (NaN, knn.cu:56); run once after launch_knn_radii + the inverse assignment
cudaError_t launch_knn_radii_fix(const uint32_t* inv_off, uint32_t K, float* radii, cudaStream_t st) {
  knn_radii_fix_kernel<<<cdivk(K, 128), 128, 0, st>>>(inv_off, K, radii);
  return cudaGetLastError();
}

// rows of the sorted order past `nv` (samples whose assignment is not a valid cluster) join the exact-search list
__global__ void knn_tail_rows_kernel(const uint32_t* __restrict__ inv, uint32_t nv, uint32_t n,
                                     uint32_t* __restrict__ rows, uint32_t* __restrict__ d_nrows) {
  uint32_t i = nv + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rows[atomicAdd(d_nrows, 1u)] = inv[i];
}
cudaError_t launch_knn_tail_rows(const uint32_t* inv, uint32_t nv, uint32_t n, uint32_t* rows, uint32_t* d_nrows,
                                 cudaStream_t st) {
  if (nv >= n) return cudaSuccess;
  knn_tail_rows_kernel<<<cdivk(n - nv, 256), 256, 0, st>>>(inv, nv, n, rows, d_nrows);
  return cudaGetLastError();
}

}  // namespace kmb

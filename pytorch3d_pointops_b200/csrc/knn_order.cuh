// Spatial ordering pre-pass for the D = 3 KNN scan (see knn_order.cu).
#pragma once
#include "common.cuh"

namespace pops {

// Everything the ordered scan reads, carved out of the caller's workspace.
struct KnnOrderBuffers {
  unsigned* maxabs_bits;  // [N]        max |coordinate| over p1 and p2 of the cloud (float bits)
  float* bbox;            // [N][6]     min xyz, max xyz of the valid p2 points
  float* blocks;          // [N][nbox][kBlockFloats]  rows x, y, z, w=|p|^2 of kBoxPoints points in curve order, the
                          //   boxes of its kSubBoxes runs of kSubPoints points, the row of original indices (u32 bits)
  float4* qsorted;        // [N][P1]    x, y, z, original index (u32 bits), curve order
  unsigned* qhome;        // [N][P1]    position in the sorted p2 where the query's code would go
  float4* boxes;          // [N][nbox][2]  (min xyz, -), (max xyz, -) of every kBoxPoints sorted p2 points
  // scratch
  unsigned* keys_in;      // [N*(P1+P2)]
  unsigned* keys_out;
  unsigned* vals_in;
  unsigned* vals_out;
  void* cub_temp;
  size_t cub_temp_bytes;
};

constexpr unsigned kNoPoint = 0xFFFFFFFFu;  // "original index" of padding entries
constexpr int kBoxPoints = 64;               // sorted p2 points per block (one bounding box each)
constexpr int kSubPoints = 16;               // consecutive sorted points per sub-box
constexpr int kSubBoxes = kBoxPoints / kSubPoints;
// One block = 1408 contiguous, 128-byte aligned bytes: rows x, y, z, w of its kBoxPoints points, the boxes of
// its kSubBoxes runs of kSubPoints points ([kSubBoxes][2] float4: (min xyz, -), (max xyz, -); runs without a
// valid point: (+inf, -inf)), then the row of original indices.  A scan needs the first kScanFloats only (one
// TMA bulk copy of 1152 bytes); the index row is read when a buffered group gets its exact distance.
constexpr int kSubOff = 4 * kBoxPoints;
constexpr int kIdxOff = kSubOff + kSubBoxes * 8;
constexpr int kScanFloats = kIdxOff;
constexpr int kBlockFloats = kIdxOff + kBoxPoints;

// boxes per cloud, padded so that a warp can read 32 boxes of any 2048-point tile in bounds
inline int64_t knn_order_num_boxes(int64_t P2) {
  const int64_t nb = (P2 + kBoxPoints - 1) / kBoxPoints;
  return (nb + 31) / 32 * 32;
}

size_t knn_order_workspace_bytes(int64_t N, int64_t P1, int64_t P2);
// Lays the buffers out inside `ws` (256-byte aligned pieces).  Returns bytes used.
size_t knn_order_carve(void* ws, int64_t N, int64_t P1, int64_t P2, KnnOrderBuffers* out);
// Runs the pre-pass on `st`.  self_knn: p1/lengths1 are the same arrays as p2/lengths2.
int knn_order_prepass(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2,
                      int N, int P1, int P2, bool self_knn, const KnnOrderBuffers& b, cudaStream_t st);

// One sort for both directions of a two-sided search: `a` = queries p1 over blocks of p2, `b` = queries
// p2 over blocks of p1 (b's scratch, maxabs_bits and bbox are not used: point them at a's).
int knn_order_prepass_pair(const float* p1, const float* p2, const int64_t* len1, const int64_t* len2, int N,
                           int P1, int P2, const KnnOrderBuffers& a, const KnnOrderBuffers& b, cudaStream_t st);

}  // namespace pops

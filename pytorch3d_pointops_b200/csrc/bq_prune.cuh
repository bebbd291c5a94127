// Ball query over Hilbert-ordered clouds (bq_prune.cu): entry point.
#pragma once
#include "knn_order.cuh"

namespace pops {

constexpr int kBqSpatialMaxK = 64;       // the hit columns hold 128 indices and are cut back to K
constexpr int kBqSpatialMinPoints = 2048;

// Decides per cloud which kernel answers it and runs the Hilbert-ordered search (bq_prune.cu) on the clouds
// it takes; the others are left untouched.  flags[n] != 0 afterwards: cloud n is answered (the index-order
// scan skips it).  `ob` is the KNN pre-pass of (p1, p2); K <= kBqSpatialMaxK.
// mode: -1 decide from sampled hit counts | 1 every finite cloud (test aid).
int bq_prune_search(const KnnOrderBuffers& ob, const float* p2, const int64_t* len1, const int64_t* len2, int N,
                    int P1, int P2, int K, float radius, float radius2, int mode, unsigned* flags, int64_t* idx,
                    float* dists, cudaStream_t st);

}  // namespace pops

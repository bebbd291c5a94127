// D = 3, L2, K <= 32 nearest neighbours over Hilbert-ordered clouds with exact bounding-box pruning.
//
// Same contract as every other KNN path (knn_cpu.cpp:13-69 of the reference: the K
// lexicographically smallest (dist, idx), dist = the unfused float32 sum), different amount of
// work: a (query, point) pair is evaluated only if the point's block can still hold a neighbour.
//
//   pre-pass (knn_order.cu)   both clouds sorted along a Hilbert curve; p2 cut into blocks of 64
//                             points (1408 contiguous bytes: rows x, y, z, w=|p|^2, the boxes of the
//                             block's four 16-point runs, the original indices) with one bounding box each.
//   this kernel               one WARP = Q*32 consecutive sorted queries, fully independent of the
//                             other warps of its CTA (no __syncthreads after start-up).
//     walk    blocks are visited outward from the warp's own position in the sorted p2.  Lane j of
//             the warp holds the lower bound between the warp's query box and block j of the
//             current 32-block chunk; one ballot against the warp's largest K-th distance selects
//             the blocks still worth reading.  The bound uses the reference's own unfused
//             operations, and rounding is monotone, so "bound > K-th distance" proves that no
//             point of the block can enter any list of the warp -- no epsilon (box_lower_bound).
//     fetch   each surviving block arrives by ONE TMA bulk copy (cp.async.bulk + mbarrier) into the
//             warp's private ring of 3 slots (K = 32: 4), up to 2 (3) blocks ahead of the scan; only the part of a
//             block a scan reads travels (rows x, y, z, w and the run boxes: 1152 bytes).
//     seed    before the first scan the warp evaluates the exact distances to its first 1-4 blocks
//             and keeps, per query, the minimum over each of 2 KT interleaved subsets; the KT-th
//             smallest of those minima bounds the K-th distance from above (KT distinct points are
//             at least that close).  The filter threshold is therefore tight from the first block
//             on: a query buffers ~1.5 K candidate points in total instead of ~K ln(P/K).
//     scan    per group of 4 points and per query: 6 FFMA2 (expanded form w - 2 q.p), min, compare,
//             predicated append of the group id to the query's candidate buffer -- as in knn.cu.
//     flush   rare (a query buffers ~K/4 + few groups in total): the buffered groups are re-read from
//             the sorted blocks in L2 (candidates hold GLOBAL group ids, so the ring can recycle
//             slots freely), exact unfused distance, 64-bit keys, register sorting networks
//             (knn_core.cuh); the lists live in the output arrays.
#include <cfloat>
#include <cstdlib>

#include "knn_prune_common.cuh"

namespace pops {

__device__ unsigned long long g_knn_stats[8];

namespace {

// Drain ONE query's candidate buffer (u32 global group ids, column stride CSTRIDE words).  Not
// inlined, called warp-converged and kept converged.  The groups are re-read from the sorted
// blocks (L2), two per step for memory-level parallelism; every point gets the exact distance and
// points with d <= dkt become 64-bit keys in the lane's survivor column; knn_merge_global folds
// them into the list.  `fresh`: the query's list is still empty and its output row unwritten (the first
// merge does not read it).  Returns (the tightened bound of the K-th distance, fresh afterwards as 0 / 1).
template <int KT, int CSTRIDE, int SSTRIDE, typename CID>
__device__ __noinline__ float2 prune_flush_one(const float* __restrict__ blocks_n, const CID* cand_col,
                                               int c_end, uint64_t* S, float q0, float q1, float q2,
                                               float dkt, bool fresh, int K, float* od, int64_t* oi) {
  constexpr unsigned FULL = 0xffffffffu;
  if (!__any_sync(FULL, c_end > 0)) return make_float2(dkt, fresh ? 1.0f : 0.0f);
  int c = 0;
  for (;;) {
    if (!__any_sync(FULL, c < c_end)) break;
    int ns = 0;
    while (c < c_end && ns <= kSurvCap - 2 * kGroup) {
      const bool two = c + 1 < c_end;
      const unsigned ga = cand_col[c * CSTRIDE];
      const unsigned gb = two ? cand_col[(c + 1) * CSTRIDE] : ga;
      c += 2;
      const float* pa = blocks_n + static_cast<size_t>(ga / kBlockGroups) * kBlockFloats + (ga % kBlockGroups) * kGroup;
      const float* pb = blocks_n + static_cast<size_t>(gb / kBlockGroups) * kBlockFloats + (gb % kBlockGroups) * kGroup;
      const float4 Xa = *reinterpret_cast<const float4*>(pa);
      const float4 Ya = *reinterpret_cast<const float4*>(pa + kBoxPoints);
      const float4 Za = *reinterpret_cast<const float4*>(pa + 2 * kBoxPoints);
      const uint4 Ia = *reinterpret_cast<const uint4*>(pa + kIdxOff);
      const float4 Xb = *reinterpret_cast<const float4*>(pb);
      const float4 Yb = *reinterpret_cast<const float4*>(pb + kBoxPoints);
      const float4 Zb = *reinterpret_cast<const float4*>(pb + 2 * kBoxPoints);
      const uint4 Ib = *reinterpret_cast<const uint4*>(pb + kIdxOff);
      float da[4], db[4];
      exact4(q0, q1, q2, Xa, Ya, Za, da);
      exact4(q0, q1, q2, Xb, Yb, Zb, db);
      const unsigned ia[4] = {Ia.x, Ia.y, Ia.z, Ia.w}, ib[4] = {Ib.x, Ib.y, Ib.z, Ib.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (da[i] <= dkt && ia[i] != kNoPoint) {
          S[ns * SSTRIDE] = make_key(da[i], ia[i]);
          ++ns;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (two && db[i] <= dkt && ib[i] != kNoPoint) {
          S[ns * SSTRIDE] = make_key(db[i], ib[i]);
          ++ns;
        }
      }
    }
    __syncwarp();
    const int ns_max = __reduce_max_sync(FULL, ns);
    if (ns_max > 0) {
      dkt = knn_merge_global<KT, SSTRIDE>(S, ns, ns_max, K, od, oi, dkt, fresh);
      fresh = fresh && ns == 0;
    }
    __syncwarp();
  }
  return make_float2(dkt, fresh ? 1.0f : 0.0f);
}

// CTAs per SM by variant (64-thread CTAs; sets the register cap): K = 1 and K = 4 fit 96 registers without
// spilling (10 CTAs; K = 4: 246 -> 228 us on the T shape), K = 8 and K = 16 stay at 8 CTAs / 128 registers
// (K = 8 at 9 CTAs / 112 registers: 331 vs 326 us; K = 16 spills below 128), K = 32 and the 4-query forms as before
constexpr int prune_min_ctas(int Q, int KT) { return KT > 16 ? 4 : (Q >= 4 ? 6 : ((Q == 1 && KT <= 4) ? 10 : 8)); }

template <int Q, int KT, int THREADS, typename CID>
__global__ void __launch_bounds__(THREADS, prune_min_ctas(Q, KT))
knn_prune_kernel(const KnnPruneParams prm) {
  constexpr int QPB = Q * THREADS, S = prune_ring_slots(KT), PF = S - 1;  // ring slots, blocks in flight ahead of the scan
  // which runs of kSubPoints points of a fetched block are scanned: every query tests the runs' boxes
  // against its own bound (K <= 16), or the warp's query box against the warp's largest bound (K = 32: the
  // extra tests cost more than the 30 % fewer runs give back, 2.36 vs 2.26 ms on the T shape; knn_subq
  // forces the per-query test)
  constexpr bool REFINE = KT <= 16;
  static_assert(kChunk * kGroup == kSubPoints && kSubBoxes == 4, "one overflow-check chunk of the scan = one sub-box");
  // seed blocks: home and its two neighbours in curve order (>= 4 seed points per tournament subset for
  // K <= 16; with one query per thread 3 blocks beat 2 on the T shape: 0.914 vs 0.955 ms at K=16)
  const int NSEED = prm.nseed > 0 ? min(prm.nseed, S) : (KT <= 16 ? 3 : 4);  // the seed blocks sit in the ring together
  using SM = PruneSmem<Q, THREADS, CID, KT>;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr unsigned FULL = 0xffffffffu;

  const int n = blockIdx.y;
  const int q_base = blockIdx.x * QPB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = prm.K;
  int64_t L1l = prm.len1[n], L2l = prm.len2[n];
  const int L1 = static_cast<int>(L1l < 0 ? 0 : (L1l > prm.P1 ? prm.P1 : L1l));
  const int L2 = static_cast<int>(L2l < 0 ? 0 : (L2l > prm.P2 ? prm.P2 : L2l));
  int64_t* out_idx = prm.idx + (static_cast<size_t>(n) * prm.P1) * K;
  float* out_d = prm.dists + (static_cast<size_t>(n) * prm.P1) * K;
  const float4* qs = prm.qsorted + static_cast<size_t>(n) * prm.P1;
  const float INF = __int_as_float(0x7f800000);
  if (prm.maxabs_bits[n] >= kDirtyBits) return;  // non-finite / huge coordinates: the exact generic kernel answers this cloud

  // CTA entirely beyond lengths1[n] (or nothing to search): rows are (0, 0).
  if (q_base >= L1 || L2 == 0) {
    const int rows = min(QPB, prm.P1 - q_base);
    for (int e = tid; e < rows * K; e += THREADS) {
      const size_t row = __float_as_uint(qs[q_base + e / K].w);
      out_idx[row * K + e % K] = 0;
      out_d[row * K + e % K] = 0.0f;
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::bars_off) + warp * S;
  float* ring = reinterpret_cast<float*>(smem + SM::ring_off) + static_cast<size_t>(warp) * S * kScanFloats;
  const float4* ring4 = reinterpret_cast<const float4*>(ring);
  CID* cand = reinterpret_cast<CID*>(smem + SM::cand_off);
  uint64_t* surv = reinterpret_cast<uint64_t*>(smem + SM::surv_off);
  // per-query state the dense loop never touches lives in shared memory, not in registers (|q|^2: qq_of)
  float* cold_dk = reinterpret_cast<float*>(smem + SM::cold_off);  // upper bound of the final K-th distance (-1: beyond lengths1)
  // output row = original query index; top bit: the list is still empty and the row unwritten
  unsigned* cold_row = reinterpret_cast<unsigned*>(cold_dk + QPB);
  constexpr unsigned kFresh = 0x80000000u;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  // ---- per-thread query state ---------------------------------------------------------------
  const float M = __uint_as_float(prm.maxabs_bits[n]);
  // E >= 130.2 * 2^-24 * M^2 bounds |filter - reference| (DESIGN.md "filter error bound").
  const float E = fmaf(M * M, 1.52587890625e-05f /* 2^-16 */, 1e-37f);
  constexpr uint32_t CB = sizeof(CID);
  constexpr uint32_t CBYTES = QPB * CB;  // bytes between consecutive entries of one candidate buffer
  const int slot0 = warp * (Q * 32) + lane;  // a warp's Q*32 queries are contiguous in curve order
  const int wq0 = q_base + warp * (Q * 32);
  float a[Q][3];   // -2 q_d (FFMA2 takes it as a broadcast scalar operand)
  float T[Q];      // filter threshold
  uint32_t cw[Q];  // shared-memory byte address of the next free candidate slot
  const uint32_t cand_base = smem_u32(cand) + static_cast<uint32_t>(slot0) * CB;
  float wqlo[3] = {INF, INF, INF}, wqhi[3] = {-INF, -INF, -INF};  // box of the warp's valid queries
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const int slot = slot0 + t * 32;
    const int qi = q_base + slot;
    const bool valid = qi < L1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qi < prm.P1) v = qs[qi];
    const unsigned row = __float_as_uint(v.w);
    const float qv[3] = {valid ? v.x : 0.0f, valid ? v.y : 0.0f, valid ? v.z : 0.0f};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      a[t][d] = -2.0f * qv[d];
      if (valid) {
        wqlo[d] = fminf(wqlo[d], qv[d]);
        wqhi[d] = fmaxf(wqhi[d], qv[d]);
      }
    }
    cold_dk[slot] = valid ? INF : -1.0f;
    cold_row[slot] = row | (valid ? kFresh : 0u);
    T[t] = valid ? FLT_MAX : -INF;
    cw[t] = cand_base + static_cast<uint32_t>(t) * (32u * CB);
    // rows beyond lengths1 are final zeros; a valid query's row is first written by its first merge
    if (qi < prm.P1 && !valid) {
      float* od = out_d + static_cast<size_t>(row) * K;
      int64_t* oi = out_idx + static_cast<size_t>(row) * K;
      for (int k = 0; k < K; ++k) {
        od[k] = 0.0f;
        oi[k] = 0;
      }
    }
  }
  // |q|^2 of query slot t as the filter uses it: a = -2 q, and scaling by powers of two commutes with rounding,
  // so this is the fmaf chain over q itself (not worth a word of shared memory per query)
  auto qq_of = [&](int t) { return 0.25f * fmaf(a[t][2], a[t][2], fmaf(a[t][1], a[t][1], a[t][0] * a[t][0])); };
  if (wq0 >= L1) return;  // no valid query in this warp; warps never meet again
  if (prm.stats && lane == 0) atomicAdd(prm.stats + 5, 1ull);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      wqlo[d] = fminf(wqlo[d], __shfl_xor_sync(FULL, wqlo[d], o));
      wqhi[d] = fmaxf(wqhi[d], __shfl_xor_sync(FULL, wqhi[d], o));
    }
  }
  // the warp's query box is only needed when a chunk of block boxes is tested (and by the K = 32 run test):
  // it waits in shared memory, not in six registers of a kernel that sits at its register cap
  float* wbox = reinterpret_cast<float*>(smem + SM::wbox_off) + warp * 8;
  if (lane < 3) {  // (selects, not wqlo[lane]: a run-time index would put the arrays in local memory)
    wbox[lane] = lane == 0 ? wqlo[0] : (lane == 1 ? wqlo[1] : wqlo[2]);
    wbox[4 + lane] = lane == 0 ? wqhi[0] : (lane == 1 ? wqhi[1] : wqhi[2]);
  }
  __syncwarp();
  float dkmax = INF;  // largest bound over the warp's valid queries

  // ---- block walk: outward from the warp's home block, 32 blocks (one chunk) per side at a time
  const int nblk = (L2 + kBoxPoints - 1) / kBoxPoints;
  const int nchunks = (nblk + 31) >> 5;
  const int nvalid_w = min(Q * 32, L1 - wq0);
  const int hb = min(static_cast<int>(prm.qhome[static_cast<size_t>(n) * prm.P1 + wq0 + (nvalid_w >> 1)]) /
                         kBoxPoints,
                     nblk - 1);
  const float4* boxes_n = prm.boxes + static_cast<size_t>(n) * prm.nbox * 2;
  auto chunk_bounds = [&](int c) -> float {  // lane j: bound for block c*32 + j
    if (!prm.prune) return 0.0f;
    const float4* bx = boxes_n + static_cast<size_t>(c * 32 + lane) * 2;  // nbox is a multiple of 32
    const float qlo[3] = {wbox[0], wbox[1], wbox[2]}, qhi[3] = {wbox[4], wbox[5], wbox[6]};
    return box_lower_bound(bx[0], bx[1], qlo, qhi);
  };
  auto valid_mask = [&](int c) -> unsigned {
    const int left = nblk - (c << 5);
    return left >= 32 ? FULL : ((1u << left) - 1u);
  };
  int chR = hb >> 5, chL = hb >> 5;
  float lbR = chunk_bounds(chR), lbL = lbR;
  unsigned pendR = valid_mask(chR) & ~((1u << (hb & 31)) - 1u);
  unsigned pendL = (1u << (hb & 31)) - 1u;
  bool go_right = true;
  float picked_lb = 0.0f;
  auto pick = [&]() -> int {  // next block whose bound does not exceed dkmax, or -1
    for (;;) {
      const bool can_r = chR < nchunks, can_l = chL >= 0;
      if (!can_r && !can_l) return -1;
      if ((go_right && can_r) || !can_l) {
        const unsigned m = pendR & __ballot_sync(FULL, lbR <= dkmax);
        if (m) {
          const int b = __ffs(m) - 1;
          pendR &= ~((2u << b) - 1u);  // skipped blocks stay pruned: dkmax only decreases
          picked_lb = __shfl_sync(FULL, lbR, b);
          go_right = false;
          return (chR << 5) + b;
        }
        if (++chR < nchunks) {
          lbR = chunk_bounds(chR);
          pendR = valid_mask(chR);
        }
      } else {
        const unsigned m = pendL & __ballot_sync(FULL, lbL <= dkmax);
        if (m) {
          const int b = 31 - __clz(m);
          pendL &= (1u << b) - 1u;
          picked_lb = __shfl_sync(FULL, lbL, b);
          go_right = true;
          return (chL << 5) + b;
        }
        if (--chL >= 0) {
          lbL = chunk_bounds(chL);
          pendL = FULL;
        }
      }
    }
  };

  // ---- ring of blocks -----------------------------------------------------------------------------
  const float* blocks_n = prm.blocks + static_cast<size_t>(n) * prm.nbox * kBlockFloats;
  float slot_lb = 0.0f;  // lane s: bound of the block in slot s
  int slot_blk = 0;      // lane s: index of the block in slot s
  int head = 0, tail = 0;  // blocks issued / scanned
  auto issue = [&](int b) {
    const int s = head % S;
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();  // the slot's previous contents were read through the generic proxy
      mbar_arrive_expect_tx(&bars[s], kBlockBytes);
      tma_bulk_g2s(ring + static_cast<size_t>(s) * kScanFloats, blocks_n + static_cast<size_t>(b) * kBlockFloats,
                   kBlockBytes, &bars[s]);
    }
    if (lane == s) {
      slot_lb = picked_lb;
      slot_blk = b;
    }
    ++head;
    if (prm.stats && lane == 0) atomicAdd(prm.stats + 0, 1ull);
  };

  // ---- flush glue ------------------------------------------------------------------------------------
  const int bufcap = prm.bufcap > 0 ? min(prm.bufcap, prune_buf_cap(KT, Q)) : prune_buf_cap(KT, Q);
  const uint32_t cw_limit = cand_base + static_cast<uint32_t>(bufcap - kChunk) * CBYTES;
  // only_full: drain just the query slots in which some lane's buffer is nearly full
  auto flush_all = [&](bool only_full) {
    if (prm.stats && lane == 0) atomicAdd(prm.stats + 2, 1ull);
    float dm = 0.0f;  // lanes beyond lengths1 carry dk = -1
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const int slot = slot0 + t * 32;
      const uint32_t base = cand_base + static_cast<uint32_t>(t) * (32u * CB);
      if (only_full && !__any_sync(FULL, cw[t] > cw_limit + static_cast<uint32_t>(t) * (32u * CB))) {
        dm = fmaxf(dm, cold_dk[slot]);
        continue;
      }
      const int c_end = static_cast<int>((cw[t] - base) / CBYTES);
      cw[t] = base;
      if (prm.stats) {
        atomicAdd(prm.stats + 3, static_cast<unsigned long long>(c_end));
        const bool anyc = __any_sync(FULL, c_end > 0);
        if (lane == 0 && anyc) atomicAdd(prm.stats + 4, 1ull);
      }
      const unsigned rw = cold_row[slot];
      const size_t row = rw & ~kFresh;
      const float2 fr = prune_flush_one<KT, QPB, THREADS, CID>(
          blocks_n, cand + slot, c_end, surv + tid, -0.5f * a[t][0], -0.5f * a[t][1], -0.5f * a[t][2],
          cold_dk[slot], (rw & kFresh) != 0u, K, out_d + row * K, out_idx + row * K);
      const float dkt = fr.x;
      if (fr.y == 0.0f) cold_row[slot] = static_cast<unsigned>(row);
      cold_dk[slot] = dkt;
      if (dkt >= 0.0f && dkt < INF) T[t] = __fadd_rn(__fsub_rn(dkt, qq_of(t)), E);
      dm = fmaxf(dm, dkt);
    }
    dkmax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(dm)));
  };

  // ---- seed: upper bound of every query's K-th distance from the first blocks ------------------
  for (int i = 0; i < NSEED; ++i) {
    const int b = pick();
    if (b < 0) break;
    issue(b);
  }
  {
    const int nseed = head;
    for (int s = 0; s < nseed; ++s) mbar_wait(&bars[s], 0);
    float dm = 0.0f;
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const int slot = slot0 + t * 32;
      const float U = seed_bound<KT>(ring4, nseed, a[t][0], a[t][1], a[t][2], qq_of(t), E);
      if (cold_dk[slot] >= 0.0f) {
        cold_dk[slot] = U;  // KT >= K distinct points lie within U
        if (U < INF) T[t] = __fadd_rn(__fsub_rn(U, qq_of(t)), E);
        dm = fmaxf(dm, U);
      }
    }
    dkmax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(dm)));
  }

  // ---- main loop ---------------------------------------------------------------------------------
  for (;;) {
    while (head - tail < PF) {
      const int b = pick();
      if (b < 0) break;
      issue(b);
    }
    if (tail == head) break;
    const int s = tail % S;
    mbar_wait(&bars[s], (tail / S) & 1);
    // dkmax may have dropped since the fetch; then the block's runs of kSubPoints points against their own
    // boxes (behind the block's rows in the ring slot): bit j of `sub` = run j is scanned
    unsigned sub = (__shfl_sync(FULL, slot_lb, s) <= dkmax) ? 0xFu : 0u;
    if (sub && prm.prune) {
      const float4* sb = ring4 + s * kBlockF4 + kSubOff / 4;
      if (REFINE || prm.subq) {
        // per query: the warp-wide test uses the box of ALL its queries against the LARGEST bound; a run is
        // scanned only if some query's own bound reaches it (same exact lower bound, with the query as a
        // degenerate box).  No state survives the test: K = 16 / 32 sit at their register cap.
        sub = 0u;
#pragma unroll
        for (int j = 0; j < kSubBoxes; ++j) {
          const float4 lo = sb[2 * j], hi = sb[2 * j + 1];
          bool need = false;
#pragma unroll
          for (int t = 0; t < Q; ++t) {
            const float qp[3] = {-0.5f * a[t][0], -0.5f * a[t][1], -0.5f * a[t][2]};
            need = need || (box_lower_bound(lo, hi, qp, qp) <= cold_dk[slot0 + t * 32]);  // dk = -1 beyond lengths1
          }
          sub |= __any_sync(FULL, need) ? (1u << j) : 0u;
        }
      } else {
        const float4 lo = sb[2 * (lane & 3)], hi = sb[2 * (lane & 3) + 1];
        const float qlo[3] = {wbox[0], wbox[1], wbox[2]}, qhi[3] = {wbox[4], wbox[5], wbox[6]};
        sub = __ballot_sync(FULL, box_lower_bound(lo, hi, qlo, qhi) <= dkmax) & 0xFu;
      }
    }
    if (sub) {
      if (prm.stats && lane == 0) {
        atomicAdd(prm.stats + 1, 1ull);
        atomicAdd(prm.stats + 6, static_cast<unsigned long long>(__popc(sub)));
      }
      const float4* tp = ring4 + s * kBlockF4;
      const unsigned gid0 = static_cast<unsigned>(__shfl_sync(FULL, slot_blk, s)) * kBlockGroups;
      // The flush is a real function call and must not sit inside the dense loop: the current and the
      // prefetched group (32 registers) and the loop's addresses would be parked in local memory around
      // it on EVERY pass (r1 profile: three LDL per chunk, 9 % of all stall samples).  On overflow the
      // loop is left, the buffers are drained, and the scan resumes at the next group with its rows
      // re-read from shared memory.
      do {
        int g = (__ffs(sub) - 1) * kChunk;  // first run still to scan
        unsigned gid = gid0 + static_cast<unsigned>(g);
        float4 Xc[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) Xc[r] = tp[r * kBlockGroups + g];
        bool over = false;
#pragma unroll 1
        for (; g < kBlockGroups && !over && ((sub >> (g / kChunk)) & 1u); g += kChunk) {
#pragma unroll
          for (int c = 0; c < kChunk; ++c) {
            // next group's rows (the last prefetch of a block reads the index row: in bounds, unused)
            float4 Xn[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) Xn[r] = tp[r * kBlockGroups + g + c + 1];
#pragma unroll
            for (int t = 0; t < Q; ++t) {
              float2 s01 = make_float2(Xc[3].x, Xc[3].y), s23 = make_float2(Xc[3].z, Xc[3].w);
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                const float2 ad = make_float2(a[t][d], a[t][d]);
                s01 = __ffma2_rn(ad, make_float2(Xc[d].x, Xc[d].y), s01);
                s23 = __ffma2_rn(ad, make_float2(Xc[d].z, Xc[d].w), s23);
              }
              const float m = fminf(fminf(s01.x, s01.y), fminf(s23.x, s23.y));
              if (m <= T[t]) {  // predicated: one STS + one IADD
                if (sizeof(CID) == 2)
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(cw[t]), "h"(static_cast<unsigned short>(gid)) : "memory");
                else
                  asm volatile("st.shared.u32 [%0], %1;" ::"r"(cw[t]), "r"(gid) : "memory");
                cw[t] += CBYTES;
              }
            }
            ++gid;
#pragma unroll
            for (int r = 0; r < 4; ++r) Xc[r] = Xn[r];
          }
          uint32_t mx = cw[0];
#pragma unroll
          for (int t = 1; t < Q; ++t) mx = max(mx, cw[t] - static_cast<uint32_t>(t) * (32u * CB));
          over = __any_sync(FULL, mx > cw_limit);
        }
        sub &= ~((1u << (g / kChunk)) - 1u);  // runs below g are done or were skipped
        if (over) flush_all(true);
      } while (sub);
    }
    ++tail;
  }
  flush_all(false);

  // the lists ARE the outputs.  Every valid query has merged at least once (its seed points pass their
  // own bound); should one not have, its row is the empty list.
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const unsigned rw = cold_row[slot0 + t * 32];
    if (rw & kFresh) {
      const size_t row = rw & ~kFresh;
      for (int k = 0; k < K; ++k) {
        out_d[row * K + k] = INF;
        out_idx[row * K + k] = static_cast<int64_t>(0xFFFFFFFFll);
      }
    }
  }
  // only slots beyond lengths2 (K > lengths2) still hold the empty marker and become the reference's
  // (0, 0) padding
  if (L2 < K) {
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const int slot = slot0 + t * 32;
      if (cold_dk[slot] < 0.0f) continue;
      const size_t row = cold_row[slot] & ~kFresh;
      for (int k = L2; k < K; ++k) {
        out_d[row * K + k] = 0.0f;
        out_idx[row * K + k] = 0;
      }
    }
  }
}

template <int Q, int KT, int THREADS, typename CID>
int launch_prune(const KnnPruneParams& prm, int N, cudaStream_t st) {
  using SM = PruneSmem<Q, THREADS, CID, KT>;
  auto kern = knn_prune_kernel<Q, KT, THREADS, CID>;
  POPS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM::total)));
  dim3 grid(static_cast<unsigned>(ceil_div(prm.P1, SM::QPB)), N);
  profile_begin("knn_scan", st);
  kern<<<grid, THREADS, SM::total, st>>>(prm);
  profile_end("knn_scan", st);
  POPS_LAUNCH_OK("knn_prune_kernel");
  return POPS_OK;
}

template <int Q, typename CID>
int launch_prune_k(const KnnPruneParams& prm, int N, cudaStream_t st) {
  constexpr int THREADS = 64;
  if (prm.K == 1) return launch_prune<Q, 1, THREADS, CID>(prm, N, st);
  if (prm.K <= 4) return launch_prune<Q, 4, THREADS, CID>(prm, N, st);
  if (prm.K <= 8) return launch_prune<Q, 8, THREADS, CID>(prm, N, st);
  if (prm.K <= 16) return launch_prune<Q, 16, THREADS, CID>(prm, N, st);
  return launch_prune<Q, 32, THREADS, CID>(prm, N, st);
}

}  // namespace

int knn_prune_search(const KnnOrderBuffers& ob, const int64_t* len1, const int64_t* len2, int N, int P1,
                     int P2, int K, int64_t* idx, float* dists, cudaStream_t st) {
  const int prune = get_option("knn_prune", 1);  // 0: visit every block (measurement aid)
  // queries per thread.  With the Hilbert order one query per thread (a warp = 32 consecutive sorted
  // queries, 128 registers, 8-10 CTAs per SM) wins on the T and chamfer shapes: K=16 0.955 vs 1.06 ms
  // (Q=4), K=4 0.41 vs 0.46, chamfer pair 0.32 vs 0.40; K=32 (218 registers at Q=1) 2.38 ms at Q=2.
  // knn_q = 1 | 2 | 4 forces one (tuning aid)
  int q = get_option("knn_q", 0);
  // (very large clouds: every warp tests all nbox / 32 chunks of boxes, which 4x more warps repeat 4x as
  //  often -- 300 K points, K=16: 0.82 ms at Q=4 vs 0.93 at Q=1; at 100 K points Q=1 is still ahead)
  if (q != 1 && q != 2 && q != 4) q = P2 >= 262144 ? 4 : (K > 16 ? 2 : 1);
  KnnPruneParams prm;
  prm.qsorted = ob.qsorted; prm.qhome = ob.qhome; prm.blocks = ob.blocks; prm.boxes = ob.boxes;
  prm.len1 = len1; prm.len2 = len2; prm.maxabs_bits = ob.maxabs_bits; prm.idx = idx; prm.dists = dists;
  prm.P1 = P1; prm.P2 = P2; prm.K = K; prm.nbox = static_cast<int>(knn_order_num_boxes(P2));
  prm.prune = prune;
  prm.nseed = get_option("knn_nseed", 0);    // seed blocks (0: 3 for K <= 16, else 4) -- tuning aid
  prm.bufcap = get_option("knn_bufcap", 0);  // candidate groups buffered per query before a flush (0: the allocated capacity)
  prm.subq = get_option("knn_subq", 0);  // 1: per-query sub-box test for every K (tuning aid; K <= 4 always)
  const int stats = get_option("knn_stats", 0);
  prm.stats = nullptr;
  if (stats) POPS_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&prm.stats), g_knn_stats));
  const bool narrow = int64_t(prm.nbox) * kBlockGroups <= 65536;
  if (q == 1) return narrow ? launch_prune_k<1, unsigned short>(prm, N, st) : launch_prune_k<1, unsigned>(prm, N, st);
  if (q == 2) return narrow ? launch_prune_k<2, unsigned short>(prm, N, st) : launch_prune_k<2, unsigned>(prm, N, st);
  return narrow ? launch_prune_k<4, unsigned short>(prm, N, st) : launch_prune_k<4, unsigned>(prm, N, st);
}

// development aid: read and reset the counters collected under POPS_KNN_STATS=1
// [0] blocks fetched, [1] blocks scanned, [2] flush rounds, [3] buffered groups, [4] non-empty
// per-slot flushes, [5] warps, [6] sub-boxes scanned
}  // namespace pops

extern "C" int pops_knn_debug_stats(unsigned long long* out8) {
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess) return POPS_ERR_CUDA;
  if (cudaMemcpyFromSymbol(out8, pops::g_knn_stats, sizeof(zero)) != cudaSuccess) return POPS_ERR_CUDA;
  if (cudaMemcpyToSymbol(pops::g_knn_stats, zero, sizeof(zero)) != cudaSuccess) return POPS_ERR_CUDA;
  return POPS_OK;
}

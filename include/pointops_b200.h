/*
 * pointops_b200.h -- C ABI of libpointops_b200.so (sm_100a CUDA kernels for the batched
 * nearest-neighbour hot path of pytorch3d_pointops).
 *
 * This is the drop-in boundary: every entry point below replaces one function the reference
 * exports from its pybind module `pytorch3d_pointops._C` (csrc/ext.cpp:15-27), or fuses a piece
 * of torch post-processing the reference does in functions/*.py.  Signatures are plain C:
 * raw DEVICE pointers, sizes, a CUDA stream handle; no torch / ATen types.  The Python host side
 * (pytorch3d_pointops_b200/_C.py) binds them with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - all tensors dense, row-major, float32 / int64 (exactly what the reference's accessors see);
 *   - every output buffer is FULLY written by the call (padding values included), so callers may
 *     pass uninitialised memory;
 *   - calls enqueue work on `stream` and return without synchronising (same contract as the
 *     reference's CUDA path, knn.cu:330-331), except where noted;
 *   - `workspace` is caller-owned scratch of at least pops_*_workspace_bytes(...) bytes, 256-byte
 *     aligned, on the same device; it may be reused by the next call on the same stream;
 *   - return value: 0 = success, otherwise a POPS_ERR_* code; pops_last_error() returns a
 *     thread-local human-readable message (the reference raises c10::Error -> RuntimeError).
 *   - inputs must be finite; |coordinate| < 1e18.  Indices are exact (ties -> lower index),
 *     distances are the reference's unfused float32 arithmetic bit for bit.
 */
#ifndef POINTOPS_B200_H_
#define POINTOPS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POPS_OK 0
#define POPS_ERR_INVALID_ARGUMENT 1 /* bad shape / norm / null pointer            */
#define POPS_ERR_WORKSPACE 2        /* workspace missing or too small             */
#define POPS_ERR_CUDA 3             /* a CUDA runtime call or kernel launch failed */
#define POPS_ERR_UNSUPPORTED 4      /* shape outside what the kernels support     */

typedef void* pops_stream_t; /* cudaStream_t */

/* Library / build identification. */
int pops_abi_version(void);
const char* pops_build_info(void); /* "sm_100a nvcc 12.9 ..." */
const char* pops_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
int64_t pops_launch_count(void);

/* Tuning / measurement knobs, also readable from the environment as POPS_<NAME> (upper case):
 *   knn_order   -1 auto | 0 never | 1 always use the curve-ordered, box-pruned D=3 search
 *   knn_prune   1 | 0: visit every block (brute force in the same order; bench.py uses it to report
 *               the evaluation rate of the scan loop next to the pruned time)
 *   knn_q       queries per thread of the pruned search (0 auto by K | 4 | 2 | 1)
 *   knn_stats   1: collect block / flush counters (pops_knn_debug_stats)
 *   knn_curve   1 Hilbert (default) | 0 Morton order in the spatial pre-pass; knn_axis_bits n: grid bits
 *               per axis of the curve codes (0 = sized to the cloud); knn_pair 1 | 0: one pre-pass for
 *               both directions in pops_knn_points_idx_pair; knn_fused_prepass 1 | 0: single-launch
 *               pre-pass (a thread-block cluster sorts a cloud, or the two clouds of a pair, in
 *               distributed shared memory) for clouds of up to 65536 points (32768 with two tensors);
 *               knn_cluster_items 0 auto | 2 | 4 | 8: sort keys per thread of that kernel
 *   knn_subq    1: every query tests the 16-point runs of a fetched block against its own bound for every K
 *               (default: K <= 16; K = 32 tests the warp's query box) -- tuning aid
 *   knn_tc      -1 auto | 0 never | 1 whenever the shape allows: tensor-core path for 32 <= D <= 256
 *   tc_cluster  CTAs that share every p2 stage by TMA multicast (1 | 2 | 4); tc_pair 1 | 0: cta_group::2 pairs
 *   tc_seed     tiles of the tensor-core scan's seed pass (default 4); tc_dbg: development switches
 *               (1 never stage a candidate, 2 drain nothing, 4 read accumulators only: NOT valid searches)
 *   bq_spatial  -1 per cloud on the device | 0 index-order scan | 1 Hilbert-ordered ball query
 *   gather_rows3 / gather_smem, knn_backward_rows  1 | 0: the specialised gather / backward kernels */
void pops_set_option(const char* name, int value);
/* Development counters of the pruned D=3 search, read and reset (needs knn_stats = 1):
 * [0] blocks fetched, [1] blocks scanned, [2] flush rounds, [3] buffered groups, [4] non-empty
 * per-slot flushes, [5] warps, [6] 16-point runs scanned. */
int pops_knn_debug_stats(unsigned long long* out8);

/* Optional per-kernel timing (bench.py "roofline"): while enabled, the library brackets its
 * dominant kernels with CUDA events on the launching stream.  pops_profile_read synchronises
 * those events and returns, for the named kernel ("knn_scan", "knn_tc_scan", "knn_tc_rerank",
 * "knn_exact_rows", "knn_generic", "fps", "ball_query", "gather", "knn_backward", "chamfer"), the number of launches seen since the last reset and their total
 * device time in milliseconds.  Off by default; costs two event records per launch when on. */
void pops_profile_enable(int on);
void pops_profile_reset(void);
int pops_profile_read(const char* kernel, int64_t* launches, double* total_ms);

/* Register-only FP32 FMA probe: runs `iters` dependent-chain FFMA rounds on every SM and returns
 * the achieved TFLOP/s (2 flop per FMA) measured with CUDA events on `stream`.  Used by bench.py
 * as the measured FP32 roofline denominator (MEASURED_PEAKS.json has no FP32 entry). */
double pops_fp32_peak_probe(int iters, pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * KNN forward.  Replaces _C.knn_points_idx (ext.cpp:21; knn.h:59-80; knn_cpu.cpp:13-69) AND the
 * sort + gather post-pass of functions/knn.py:77-89: results come back already in the canonical
 * order -- the K lexicographically smallest (dist, idx), ascending.
 *   p1 (N,P1,D) f32, p2 (N,P2,D) f32, lengths1/lengths2 (N) i64 (values clamped to [0,P]),
 *   norm 1|2, version: accepted for API compatibility (knn.py:121) and ignored.
 *   idx (N,P1,K) i64, dists (N,P1,K) f32: rows >= lengths1[n] and slots >= min(K,lengths2[n])
 *   are (0, 0.0f) (knn_cpu.cpp:25-26).
 * ------------------------------------------------------------------------------------------- */
size_t pops_knn_workspace_bytes(int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K, int norm);
int pops_knn_points_idx(const float* p1, const float* p2, const int64_t* lengths1,
                        const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2, int64_t D,
                        int64_t K, int norm, int version, int64_t* idx, float* dists,
                        void* workspace, size_t workspace_bytes, pops_stream_t stream);

/* The same search in two phases, for host pipelines that want the result slice by slice (the outer
 * `for n` of knn_cpu.cpp:35 is independent per cloud; host.py: HostKnn overlaps the device-to-host
 * copy of one slice with the search of the next).  Additive: the reference has no counterpart.
 *   pops_knn_points_prepare     everything that precedes the search, once for ALL N clouds (the
 *                               spatial pre-pass of the D = 3 path; nothing for the other paths)
 *   pops_knn_points_idx_range   the rows of clouds [n0, n1) of idx / dists, which are the FULL
 *                               (N,P1,K) arrays; all other arguments as passed to prepare
 * Same stream (or ordered after prepare), same workspace of pops_knn_workspace_bytes(N, ...). */
int pops_knn_points_prepare(const float* p1, const float* p2, const int64_t* lengths1,
                            const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2, int64_t D,
                            int64_t K, int norm, void* workspace, size_t workspace_bytes,
                            pops_stream_t stream);
int pops_knn_points_idx_range(const float* p1, const float* p2, const int64_t* lengths1,
                              const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2, int64_t D,
                              int64_t K, int norm, int version, int64_t n0, int64_t n1, int64_t* idx,
                              float* dists, void* workspace, size_t workspace_bytes,
                              pops_stream_t stream);

/* Both directions of a two-sided search in one call -- what chamfer_distance needs
 * (functions/chamfer.py:136 calls knn_points(x, y) and, for the other direction, knn_points(y, x)):
 * idx12/dists12 (N,P1,K) = neighbours of p1 in p2, idx21/dists21 (N,P2,K) = neighbours of p2 in p1,
 * each exactly what pops_knn_points_idx returns for that direction.  On the D = 3 path both clouds
 * are ordered by ONE pre-pass (one sort instead of two).  Additive. */
size_t pops_knn_pair_workspace_bytes(int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K, int norm);
int pops_knn_points_idx_pair(const float* p1, const float* p2, const int64_t* lengths1,
                             const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2, int64_t D,
                             int64_t K, int norm, int64_t* idx12, float* dists12, int64_t* idx21,
                             float* dists21, void* workspace, size_t workspace_bytes,
                             pops_stream_t stream);

/* _C.knn_check_version (ext.cpp:19; knn.cu:292-303): which (D,K) the reference's kernel
 * variant `version` accepts.  Kept so callers probing it keep working; returns 0/1. */
int pops_knn_check_version(int version, int64_t D, int64_t K);

/* ---------------------------------------------------------------------------------------------
 * KNN backward.  Replaces _C.knn_points_backward (ext.cpp:22; knn.h:127-149; knn_cpu.cpp:75-128).
 *   grad_p1[n,i,:] += sum_k diff, grad_p2[n,idx,:] -= diff, diff = 2*g*(p1-p2) (L2) or g*sign (L1);
 *   k < min(K, lengths2[n]), i < lengths1[n], idx == -1 skipped (ball_query reuses this, :49-51).
 *   grad_p1 (N,P1,D) and grad_p2 (N,P2,D) are zero-filled by the call.  float atomics on grad_p2.
 * ------------------------------------------------------------------------------------------- */
int pops_knn_points_backward(const float* p1, const float* p2, const int64_t* lengths1,
                             const int64_t* lengths2, const int64_t* idx, const float* grad_dists,
                             int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K, int norm,
                             float* grad_p1, float* grad_p2, pops_stream_t stream);
/* The same with caller-owned scratch of pops_knn_backward_workspace_bytes(N, P2, D) bytes (additive):
 * for D = 3 the grad_p2 scatter then runs as one 16-byte vector reduction per (query, neighbour)
 * into a float4-padded copy of grad_p2 that a second small kernel folds into the 12-byte rows -- a
 * third of the L2 atomic operations of three scalar reductions.  workspace NULL = the call above. */
size_t pops_knn_backward_workspace_bytes(int64_t N, int64_t P2, int64_t D);
int pops_knn_points_backward_ws(const float* p1, const float* p2, const int64_t* lengths1,
                                const int64_t* lengths2, const int64_t* idx, const float* grad_dists,
                                int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K, int norm,
                                float* grad_p1, float* grad_p2, void* workspace,
                                size_t workspace_bytes, pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Ball query.  Replaces _C.ball_query (ext.cpp:23; ball_query.h:62-93; ball_query_cpu.cpp:12-54):
 * the first K points of p2 (index order) with dist2 < radius*radius (strict, f32 product).
 *   idx (N,P1,K) i64 padded with -1, dists (N,P1,K) f32 padded with 0.
 * ------------------------------------------------------------------------------------------- */
size_t pops_ball_query_workspace_bytes(int64_t N, int64_t P1, int64_t P2, int64_t D, int64_t K);
int pops_ball_query(const float* p1, const float* p2, const int64_t* lengths1,
                    const int64_t* lengths2, int64_t N, int64_t P1, int64_t P2, int64_t D,
                    int64_t K, float radius, int64_t* idx, float* dists, void* workspace,
                    size_t workspace_bytes, pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Farthest point sampling.  Replaces _C.sample_farthest_points (ext.cpp:24;
 * sample_farthest_points.h:55-76; sample_farthest_points_cpu.cpp:14-103).
 *   points (N,P,D) f32, lengths/K/start_idxs (N) i64; idx (N,max_K) i64 padded with -1,
 *   idx[n,0] = start_idxs[n]; ties -> lowest index.  max_K is supplied by the caller (the
 *   reference syncs on torch::max(K), sample_farthest_points.cu:132); no device sync here.
 * ------------------------------------------------------------------------------------------- */
size_t pops_fps_workspace_bytes(int64_t N, int64_t P, int64_t D, int64_t max_K);
int pops_sample_farthest_points(const float* points, const int64_t* lengths, const int64_t* K,
                                const int64_t* start_idxs, int64_t N, int64_t P, int64_t D,
                                int64_t max_K, int64_t* idx, void* workspace,
                                size_t workspace_bytes, pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Ragged copies.  Replace _C.packed_to_padded / _C.padded_to_packed (ext.cpp:16-17;
 * packed_to_padded_tensor.h:78-113; packed_to_padded_tensor_cpu.cpp:11-70).
 *   packed (F,D) f32, first_idxs (B) i64, padded (B,max_size,D) f32; cloud b owns rows
 *   [first_idxs[b], b+1<B ? first_idxs[b+1] : F).  Outputs are zero where nothing is copied.
 * ------------------------------------------------------------------------------------------- */
int pops_packed_to_padded(const float* packed, const int64_t* first_idxs, int64_t num_inputs,
                          int64_t B, int64_t max_size, int64_t D, float* padded,
                          pops_stream_t stream);
int pops_padded_to_packed(const float* padded, const int64_t* first_idxs, int64_t num_inputs,
                          int64_t B, int64_t max_size, int64_t D, float* packed,
                          pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * sample_pdf.  Replaces _C.sample_pdf (ext.cpp:27; sample_pdf.h:58-78; sample_pdf_cpu.cpp:24-142):
 * inverse-CDF sampling of B piecewise-constant densities.  bins (B,n_bins+1) f32 bin edges,
 * weights (B,n_bins) f32 >= 0, outputs (B,n_samples) f32: on entry uniform numbers in [0,1], on
 * return the samples -- IN PLACE, like the reference (which also bumps the tensor's autograd
 * version; the Python glue does that).  Same float operations as the reference CPU path.
 * ------------------------------------------------------------------------------------------- */
int pops_sample_pdf(const float* bins, const float* weights, float* outputs, int64_t B,
                    int64_t n_bins, int64_t n_samples, float eps, pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused row gather.  Replaces the torch expand+gather+mask of knn_gather (functions/knn.py:200-250)
 * and masked_gather (functions/utils.py:20-65).
 *   x (N,M,U) f32, idx (N,L,K) i64 -> out (N,L,K,U) f32.
 *   mode POPS_GATHER_KNN:    out[n,l,k] = x[n, idx[n,l,k]], zero where k >= lengths[n]
 *                            (lengths may be NULL = no masking); an index outside [0,M) is an
 *                            error in the reference (RuntimeError) -- here it sets *oob_flag = 1
 *                            (int32 on device, may be NULL) and yields zeros.
 *   mode POPS_GATHER_MASKED: idx == -1 -> zero row.
 * pops_gather_backward scatters grad_out (N,L,K,U) into grad_x (N,M,U) (zero-filled by the call).
 * ------------------------------------------------------------------------------------------- */
#define POPS_GATHER_KNN 0
#define POPS_GATHER_MASKED 1
int pops_gather(const float* x, const int64_t* idx, const int64_t* lengths, int64_t N, int64_t M,
                int64_t U, int64_t L, int64_t K, int mode, float* out, int32_t* oob_flag,
                pops_stream_t stream);
int pops_gather_backward(const float* grad_out, const int64_t* idx, const int64_t* lengths,
                         int64_t N, int64_t M, int64_t U, int64_t L, int64_t K, int mode,
                         float* grad_x, pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused neighbourhood covariances (additive).  Replaces the torch sequence of get_point_covariances
 * (functions/utils.py:111-153) after the KNN search: gather, mean over the K slots, centring, outer
 * products, mean.  x (N,M,D) f32 with 1 <= D <= 4, idx (N,P,K) i64 from pops_knn_points_idx,
 * lengths (N) i64 or NULL (slots k >= lengths[n] gather zeros, as knn_gather does)
 * -> nn (N,P,K,D), cov (N,P,D,D).
 * ------------------------------------------------------------------------------------------- */
int pops_point_covariances(const float* x, const int64_t* idx, const int64_t* lengths, int64_t N,
                           int64_t P, int64_t M, int64_t D, int64_t K, float* nn, float* cov,
                           pops_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused chamfer post-processing (additive).  Replaces the ~20 torch kernels per direction that
 * functions/chamfer.py:114-189 runs after the K=1 search (mask, weights, knn_gather,
 * cosine_similarity, abs, 1-x, point reduction) and their autograd mirror images.
 *   dists (N,P1) f32 / idx (N,P1) i64: output of pops_knn_points_idx with K = 1;
 *   weights (N) f32 or NULL;  xf[f] (N,P1,chans[f]) / yf[f] (N,P2,chans[f]): num_feats <= 8 feature
 *   pairs (host arrays of device pointers);  point_reduction 0 none | 1 sum | 2 mean | 3 max.
 *   forward: cham_out (N) [or (N,P1) for none], feat_out (F,N) [or (F,N,P1)], argmax_out (N) for max.
 *   backward: g_cham / g_feat shaped like the forward outputs; grad_x (N,P1,D), grad_y (N,P2,D),
 *   grad_xf[f], grad_yf[f] are zero-filled and written by the call (float atomics on the y side);
 *   accumulate != 0: the buffers are NOT cleared and the call adds to them -- the y -> x direction of
 *   a two-sided loss lands on the x -> y direction's gradients without a separate sum.
 *   g_broadcast != 0 (point_reduction sum / mean only): g_cham is ONE scalar and g_feat holds one
 *   scalar per feature, shared by all N clouds (the gradient of a batch-reduced loss); g_scale
 *   multiplies them (1/N for a batch mean, else 1).
 * ------------------------------------------------------------------------------------------- */
int pops_chamfer_forward(const float* dists, const int64_t* idx, const int64_t* lengths1,
                         const int64_t* lengths2, const float* weights, int64_t N, int64_t P1,
                         int64_t P2, int num_feats, const float* const* xf, const float* const* yf,
                         const int64_t* chans, int point_reduction, int abs_cosine, float* cham_out,
                         float* feat_out, int64_t* argmax_out, pops_stream_t stream);
int pops_chamfer_backward(const float* x, const float* y, const int64_t* idx, const int64_t* lengths1,
                          const int64_t* lengths2, const float* weights, int64_t N, int64_t P1,
                          int64_t P2, int64_t D, int norm, int num_feats, const float* const* xf,
                          const float* const* yf, const int64_t* chans, int point_reduction,
                          int abs_cosine, const float* g_cham, const float* g_feat,
                          const int64_t* argmax, float* grad_x, float* grad_y, float* const* grad_xf,
                          float* const* grad_yf, int accumulate, int g_broadcast, float g_scale,
                          pops_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POINTOPS_B200_H_ */

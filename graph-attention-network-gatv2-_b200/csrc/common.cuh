// Shared device helpers and launcher declarations of libgatx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gatx {

constexpr float kSlope = 0.01f;  // LeakyReLU slope of the reference (EB:1143, EB:1428): the default of both slopes
constexpr int kNumSMs = 148;     // B200

// The two LeakyReLU slopes of a layer: `attn` inside the attention score a . LReLU(W_l x_j + W_r x_i) (EB:1143; GATv2's
// negative_slope) and `act` of the layer activation (EB:1428).  The reference fixes both at 0.01; 0 <= slope < 1.
struct Slopes {
  float attn, act;
};
__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : slope * x; }
__device__ __forceinline__ float lrelu_grad(float x, float slope) { return x > 0.f ? 1.f : slope; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming 128-bit load that does not allocate in L1 (gathered rows are used once per warp)
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- launchers (each returns the number of kernels it launched) --------------------------

// graph_prep.cu
int launch_csr_to_coo(const int* row_ptr, const int* col_idx, int* src, int* dst, int* deg, int n_rows,
                      cudaStream_t st);
// builds the source-major transpose of a (local) CSR: csc_ptr[n_src+1], csc_dst[E], csc_eid[E];
// stable (ascending CSR position inside every source segment).  Returns launches, <0 on error.
int build_csc(const int* col_idx, const int* coo_dst, int64_t E, int n_src, int* csc_ptr, int* csc_dst,
              int* csc_eid, cudaStream_t st);
int launch_philox_uniform(float* out, int64_t n, float limit, uint64_t seed, uint64_t stream_id,
                          cudaStream_t st);
// Y[n][c] = keep(n, c) ? X[n][c] / (1 - p) : 0 for rows [row0, row0 + n_rows) of a layer input (Y == X: in place);
// keep() is a pure function of (seed, layer, step, global row, column), see dropout_kernel
int launch_dropout(const float* X, int64_t ldx, float* Y, int64_t ldy, int n_rows, int cols, int row0, float p,
                   uint64_t seed, int layer, int64_t step, cudaStream_t st);
// out[e][h] = keep ? 1 / (1 - p) : 0 for local edges [0, E): (e, h) is kept iff word h % 4 of
// Philox4x32-10(counter {edge0 + e, h / 4, 2^31 | layer, step}, key seed) >= floor(p * 2^32)  (edge0 + e: GLOBAL CSR position)
int launch_attn_dropout_scale(float* out, int64_t E, int H, int64_t edge0, float p, uint64_t seed, int layer, int64_t step,
                              cudaStream_t st);

// gemm_simt.cu : C[m][n] (ldc) (+)= sum_k A(m,k) B(n,k), A(m,k) = A[m*sAm + k*sAk], same for B.
int launch_gemm_simt(const float* A, int64_t sAm, int64_t sAk, const float* B, int64_t sBn, int64_t sBk,
                     float* C, int64_t ldc, int M, int N, int64_t K, bool accumulate, float* splitk_ws,
                     size_t splitk_ws_bytes, cudaStream_t st);
// 3xTF32 operand split: hi = x with the low 13 mantissa bits cleared (exactly what kind::tf32 reads), lo = x - hi (exact
// in fp32).  Writes hi to out_hi[r][c] and lo to out_lo[r][c] for r < rows, c < cols; columns [cols, cols_pad) are zeroed.
int launch_split_tf32(const float* X, int64_t ldx, int64_t rows, int cols, int cols_pad, float* out_hi, int64_t ld_hi,
                      float* out_lo, int64_t ld_lo, cudaStream_t st);
// packs EB-layout W [F][2I] into Wcat [2F][ldk] (rows 0..F-1 = W_l, F..2F-1 = W_r, zero padded)
// and WcatT [I][2F]
int launch_pack_weights(const float* W, int F, int I, float* Wcat, int ldk, float* WcatT, cudaStream_t st);

// edge_kernels.cu
struct EdgeGraph {
  int n_rows;            // local destination rows
  const int* row_ptr;    // [n_rows+1] local CSR (positions are local edge ids)
  const int* col_idx;    // [E] global source ids
  int n_src;             // number of source rows (global N)
  const int* csc_ptr;    // [n_src+1]
  const int* csc_dst;    // [E] local destination row
  const int* csc_eid;    // [E] local edge id
  const int* heavy_rows; int n_heavy_rows;   // destination rows with degree > kHeavyDeg
  const int* heavy_srcs; int n_heavy_srcs;   // sources with out-degree > kHeavyDeg
  int64_t E;
  // edge-balanced chunks (edge_stream.cu): chunk c = edges [c*chunk_T, (c+1)*chunk_T)
  int chunk_T, n_chunks;
  const int* chunk_row;  // [n_chunks] destination row containing edge c*chunk_T
  const int* chunk_src;  // [n_chunks] source row containing transposed position c*chunk_T
  // L2 residency hints for the gathers (nullptr = off): copies of col_idx / csc_dst whose two top bits mark the
  // nodes gathered most often (bit 31: the hot set sized for rows of hot_wide_F floats, bit 30: for narrower rows)
  const int* col_idx_hot;
  const int* csc_dst_hot;
  int hot_wide_F;
  // optional CUDA-event pairs around the three main streaming kernels of the layer being launched
  // (fwd, bwd pass 1, bwd pass 2); null = no timing
  cudaEvent_t* kernel_events;  // [6] = {fwd_a, fwd_b, dst_a, dst_b, src_a, src_b}
  int reserve_ctas;            // CTA slots the streaming kernels leave to the exchange kernels running underneath (0 = none)
  Slopes slopes;               // LeakyReLU slopes of the layer being launched
  const float* bias;           // [F] added to the aggregate before the activation (extension); nullptr = none (reference)
  // attention-coefficient dropout (extension; nullptr = off, the reference): [E][H] keep / (1 - p) per (local edge, head).
  // Forward: h = sum alpha * ascale * P_l (the softmax itself is not touched); pass 1: galpha carries ascale and the
  // record keeps alpha * ascale for pass 2.
  const float* ascale;
};
constexpr int kHeavyDeg = 1024;
bool edge_shape_supported(int H, int D);
int edge_rec_words(int H, int D);  // 32-bit words per edge of the backward record
int launch_edge_forward(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                        float* Hout, float* hpre /*nullable*/, float* score, float* mx, float* sinv,
                        cudaStream_t st);
// pass 1 (destination-major): gH (post-activation grad, [n_rows][F]) is overwritten in place with the
// pre-activation grad; writes gPr, the per-edge record and per-block partial sums of ga.
int launch_edge_backward_dst(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr,
                             const float* a, const float* Hout, float* gH, const float* score,
                             const float* mx, const float* sinv, float* gPr, uint32_t* rec,
                             float* ga_partials, int* n_partials, float* galpha_dbg, cudaStream_t st);
// pass 2 (source-major): gPl[n_src][F]
int launch_edge_backward_src(const EdgeGraph& g, int H, int D, const float* a, const float* gH,
                             const uint32_t* rec, float* gPl, cudaStream_t st);
constexpr int kEdgeBwdBlocks = kNumSMs * 4;
// out[i] (+)= sum_b partials[b][i], fixed order
int launch_reduce_partials(const float* partials, int n_partials, int n, float* out, bool accumulate,
                           cudaStream_t st);
int launch_unpack_rec(const uint32_t* rec, int64_t E, int H, int D, float* alpha, float* ge, cudaStream_t st);
int launch_alpha_from_score(const float* score, const int* coo_dst, const float* mx, const float* sinv, int64_t E,
                            int H, float* alpha, cudaStream_t st);
// out[c] (+)= sum over rows of M[row][c] (fixed order: per-block partials in `partials`, >= kColsumBlocks * cols floats)
constexpr int kColsumBlocks = kNumSMs * 2;
int launch_colsum(const float* M, int n_rows, int cols, float* partials, float* out, bool accumulate, cudaStream_t st);
int launch_head_mean(const float* Hfull, int N, int H, int D, float* Hout, cudaStream_t st);
int launch_head_bcast_grad(const float* gHout, int N, int H, int D, float* gHfull, cudaStream_t st);

// edge_stream.cu : edge-balanced streaming kernels (bulk-async-copy ring) for rows of 128/256/512 floats
bool edge_stream_supported(int H, int D);
int64_t edge_stream_part_floats(int H, int D, int n_chunks);
int launch_chunk_rows(const int* row_ptr, int n_rows, int64_t E, int T, int n_chunks, int* chunk_row, cudaStream_t st);
int launch_edge_forward_stream(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                               float* Hout, float* hpre, float* score, float* mx, float* sinv, float* part,
                               cudaStream_t st);
// phases bit 0: prep (g_h in place + cdot) + pass 1 (gPr, rec, ga partials) over g's destination rows;
// bit 1: pass 2 (gPl) over g's sources; bit 2 (with bit 0): skip the prep, g_h already holds the pre-activation gradient
// and cdot the segment sums (written by the input-gradient GEMM's fused epilogue, gemm_tc.cuh GemmFuse).  The multi-GPU epoch runs pass 1 per block of destination rows (a block is
// an EdgeGraph of its own: rebased row_ptr, its own chunks, pointers offset to the block) and pass 2 once.
int launch_edge_backward_stream(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                                const float* Hout, float* gH, float* cdot, const float* score, const float* mx,
                                const float* sinv, float* gPr, float* gPl, uint32_t* rec, float* part,
                                float* ga_partials, int* n_partials, float* galpha_dbg, cudaStream_t st, int phases = 3);

// edge_generic.cu : any heads <= 32, any per-head dim with heads*dim <= 1024 (scalar fallback)
bool edge_generic_supported(int H, int D);
int edge_generic_rec_words(int H);
int edge_generic_partials();
int launch_edge_forward_generic(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                                float* Hout, float* hpre, float* score, float* mx, float* sinv, cudaStream_t st);
int launch_edge_backward_generic(const EdgeGraph& g, int H, int D, const float* Pl, const float* Pr, const float* a,
                                 const float* Hout, float* gH, const float* score, const float* mx, const float* sinv,
                                 float* gPr, float* gPl, uint32_t* rec, float* ga_partials, int* n_partials,
                                 float* galpha_dbg, cudaStream_t st);
int launch_unpack_rec_generic(const uint32_t* rec, int64_t E, int H, float* alpha, float* ge, cudaStream_t st);

// head_loss.cu : classifier + softmax + CE + argmax + output gradients (EB:463-608)
constexpr int kHeadBlocks = kNumSMs * 2;
int launch_head(const float* HL, const float* Wo, const int* labels, int N, int C, int DL, int ldc, float* y, float* dz,
                float* z_dbg, int* pred, float* gH, double* loss_partials, int* correct_partials, int* n_partials,
                const unsigned char* mask, cudaStream_t st);
// tensor-core head: z = H_L W_o^T and gH = dz W_o run as tcgen05 GEMMs around this kernel
int launch_softmax_ce(const float* z, const int* labels, int N, int C, int ldc, float* y, float* dz, int* pred,
                      double* loss_partials, int* correct_partials, int* n_partials, const unsigned char* mask,
                      cudaStream_t st);
int launch_transpose_wo(const float* Wo, int C, int DL, int ldc, float* WoT, cudaStream_t st);
int launch_loss_finalize(const double* loss_partials, const int* correct_partials, int n_partials,
                         double* loss_sum, long long* correct, cudaStream_t st);

// halo_p2p.cu : halo exchange through NVLink peer memory (destination-row partition, world > 1)
constexpr int kMaxPeers = 16;
struct PeerPtrs {
  float* p[kMaxPeers];  // base of the peer's [N][F] buffer (this rank's own slot is unused)
};
// max_ctas > 0 caps the grid (the pipelined exchange runs underneath the edge passes in a fixed number of CTA slots)
int launch_halo_push(const float* own_rows, int r0, int n_rows, int F, const uint16_t* ref_mask, const PeerPtrs& peers,
                     int me, cudaStream_t st, int max_ctas = 0);
// backward exchange: scatter of partial rows into the owners' staging buffers, then the owner's local ordered sum
struct ScatterPlan {
  int n_seg;                    // one segment per other rank
  int cum[kMaxPeers + 1];       // prefix sums of the segments' row counts
  int row0[kMaxPeers];          // first global row of the segment (a row block of its owner)
  int owner_row0[kMaxPeers];    // first global row the owner owns
  int owner[kMaxPeers];         // rank that owns the segment's rows
  float* dst[kMaxPeers];        // owner's staging slot of THIS rank: [owner's n_rows][F]
};
// bulk-copy (TMA) transport (GATX_HALO_MODE=bulk): one lane per warp moves whole rows global -> shared ring -> peer global
bool halo_bulk_supported(int F);
int launch_halo_push_bulk(const float* own_rows, int r0, int n_rows, int F, const uint16_t* ref_mask, const PeerPtrs& peers,
                          int me, cudaStream_t st, int max_ctas);
int launch_halo_scatter_bulk(const float* partial, int F, const unsigned char* my_ref, const ScatterPlan& plan,
                             const PeerPtrs& stage_base, int me, cudaStream_t st, int max_ctas);
int launch_halo_sum(float* own_rows, int row_off, int n_rows, int n_rows_total, int F, const uint16_t* ref_mask,
                    const float* stage, int me, int world, cudaStream_t st);
int launch_halo_pull(float* own_rows, int r0, int n_rows, int F, const uint16_t* ref_mask, const PeerPtrs& peers, int me,
                     int world, cudaStream_t st, int max_ctas = 0);
int halo_cta_slots(int world);  // CTA slots of the exchange kernels (GATX_HALO_CTAS); 0 = no cap, no reservation (default)
// Device-side barrier across ranks through flags in peer memory: rank `me` stores `seq` (release, system scope) into
// slot `me` of every peer's flag array, then waits (acquire) until every slot of its own array has reached `seq`.
// Stream-ordered: everything this rank enqueued before it on `st` (and its peer-memory stores) is visible to a peer
// once that peer has left its own barrier `seq`.  A peer that never arrives trips a timeout (the kernel traps).
struct PeerFlags {
  uint32_t* p[kMaxPeers];  // flag array [kMaxPeers] of every rank (own slot: this rank's array)
};
int launch_halo_barrier(const PeerFlags& flags, int me, int world, uint32_t seq, cudaStream_t st);

// graph_prep.cu: out[e] = idx[e] | bit31 if degree(idx[e]) >= thr_wide | bit30 if degree(idx[e]) >= thr_narrow
int launch_mark_hot(const int* idx, const int* ptr, int64_t E, int thr_wide, int thr_narrow, int* out, cudaStream_t st);

// optim.cu
struct OptimGroups {
  int64_t begin[3], end[3];  // {all W}, {all a}, {W_o} ranges of the flat parameter buffer
};
constexpr int kOptimBlocks = 256;
int launch_optimizer(float* params, float* grads, float* m, float* v, int64_t n, OptimGroups grp, bool clip,
                     int optimizer, float lr, float b1, float b2, int t, float* norm_partials,
                     cudaStream_t st);

}  // namespace gatx

/*
 * gatx.h -- C ABI of the B200-native GATv2 full-batch training engine.
 *
 * The reference (saurabh260918/Graph-Attention-Network-GATv2-) has no operator / plugin / FFI
 * boundary: `main` launches its kernels directly (GATv2_edge_based.cu:927-1646, "EB").  This
 * header IS the boundary for the one hot path -- one full-batch training epoch
 * (EB:1370-1642) -- and every entry point names the reference code it replaces.  The library
 * behind it (libgatx.so) is plain CUDA C++ for sm_100a: no PyTorch, no Triton, no CPU
 * fallback.  If no CUDA device is usable, gatx_create fails with GATX_ERR_CUDA.
 *
 * Conventions: every call returns 0 on success or a negative gatx_status; the text of the last
 * error of a context is gatx_last_error(ctx).  Nothing throws, nothing exits.  The caller owns
 * all host buffers (copied on set_*, filled on get_*); the context owns device memory and its
 * streams.  A context is not thread-safe; distinct contexts (one per device / rank) may be
 * driven from distinct host threads or processes.
 *
 * Layouts are the reference's: CSR is destination-major (row = destination, col_idx = source,
 * EB:67-84); W of a layer is [H][D][2*I] row-major with columns 0..I-1 applied to the SOURCE
 * node and I..2I-1 to the DESTINATION node (EB:304-316); a is [H][D]; W_o is [C][D_last]
 * (EB:1243-1258).  Per-edge tensors are returned edge-major [E][H] (the reference stores
 * [H][E], EB:297); callers transpose if they need the reference order.
 */
#ifndef GATX_H_
#define GATX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gatx_ctx gatx_ctx;

typedef enum gatx_status {
  GATX_OK = 0,
  GATX_ERR_INVALID = -1,     /* bad argument / call order */
  GATX_ERR_CUDA = -2,        /* CUDA runtime error (text in gatx_last_error) */
  GATX_ERR_UNSUPPORTED = -3, /* shape outside the implemented kernels */
  GATX_ERR_NCCL = -4,
  GATX_ERR_OOM = -5
} gatx_status;

enum { GATX_OPT_SGD = 0, GATX_OPT_ADAM = 1 };
/* Projection / gradient GEMM arithmetic. */
enum {
  GATX_GEMM_TF32_TC = 0, /* tcgen05 tensor cores, TF32 inputs, fp32 accumulate (default) */
  GATX_GEMM_FP32_SIMT = 1, /* fp32 FMA on CUDA cores (tight-tolerance parity runs) */
  /* fp32-grade results ON the tensor cores: every operand is split into hi (its TF32 part) + lo (the exact remainder)
   * and hi*hi + lo*hi + hi*lo is accumulated into the same TMEM tile (the lo*lo term, 2^-22 relative, is dropped).
   * Same tolerances and bit-exact predicted labels as GATX_GEMM_FP32_SIMT at a fraction of its time. */
  GATX_GEMM_3XTF32_TC = 2
};

/* Replaces the CLI-derived locals of EB:934-1040 (L, head[], out_dim[], optimizer, lr, betas,
 * clip).  Gradient clipping threshold is the reference's fixed 5.0 (EB:1563). */
typedef struct gatx_config {
  int32_t num_layers;
  const int32_t* heads;   /* [num_layers] */
  const int32_t* outdims; /* [num_layers] per-head output dim */
  int32_t optimizer;      /* GATX_OPT_* */
  float lr, beta1, beta2;
  int32_t clip;        /* 0/1 */
  int32_t device;      /* CUDA device ordinal */
  int32_t gemm_mode;   /* GATX_GEMM_* */
  int32_t keep_debug;  /* 1: also keep pre-activation h, galpha for gatx_get_tensor */
  int32_t rank, world; /* destination-row partition; world = 1 for single GPU */
} gatx_config;

/* Tensor ids for gatx_get_tensor / gatx_tensor_size (same ids as the oracle's). */
enum {
  GATX_T_W = 0,      /* [H][D][2I]  layer parameter (EB d_w + w_offset[l]) */
  GATX_T_A = 1,      /* [H][D]      (EB d_a + a_offset[l]) */
  GATX_T_WO = 2,     /* [C][D_last] (EB d_wo) */
  GATX_T_GW = 3,     /* grad of W   (EB grad_d_w) -- valid after gatx_backward, before gatx_step */
  GATX_T_GA = 4,     /* grad of a   (EB grad_d_a) */
  GATX_T_GWO = 5,    /* grad of W_o (EB grad_wo) */
  GATX_T_PL = 6,     /* [N][F] projected source features  W_l x */
  GATX_T_PR = 7,     /* [N_local][F] projected destination features W_r x */
  GATX_T_SCORE = 8,  /* [E][H] attention logits (EB attn_score, transposed) */
  GATX_T_ALPHA = 9,  /* [E][H] attention coefficients (EB attn_coeff, transposed) */
  GATX_T_HPRE = 10,  /* [N][F] pre-activation aggregate (EB d_h) -- needs keep_debug */
  GATX_T_HOUT = 11,  /* [N][F] or [N][D_last] layer output (EB d_layer_outputs) */
  GATX_T_Y = 12,     /* [N][C] class probabilities (EB d_y) */
  GATX_T_GH = 13,    /* [N][F] grad wrt pre-activation aggregate (EB input_gradients[l]) */
  GATX_T_Z = 14,     /* [N][C] logits before softmax -- needs keep_debug */
  GATX_T_PRED = 15,  /* int32 [N] predicted labels (EB:530-535 first-max argmax) */
  GATX_T_COO_SRC = 16, /* int32 [E] (EB d_src, EB:67-84) */
  GATX_T_COO_DST = 17, /* int32 [E] (EB d_dst) */
  GATX_T_IN_DEGREE = 18, /* int32 [N] row lengths (EB:93) */
  GATX_T_CSC_PTR = 19,   /* int32 [N+1] transposed graph used by the deterministic backward */
  GATX_T_CSC_DST = 20,   /* int32 [E] */
  GATX_T_CSC_EID = 21,   /* int32 [E] CSR position of each transposed edge */
  GATX_T_GPL = 22,   /* [N][F] grad wrt projected source features */
  GATX_T_GPR = 23,   /* [N][F] grad wrt projected destination features */
  GATX_T_GALPHA = 24, /* [E][H] grad wrt attention coefficients (EB grad_attn_coeff, transposed) -- keep_debug */
  GATX_T_GE = 25,    /* [E][H] grad wrt attention logits (EB grad_attn_score, transposed) -- keep_debug */
  GATX_T_B = 26,     /* [H*D] layer bias (extension, gatx_set_bias) */
  GATX_T_GB = 27     /* grad of the layer bias */
};

/* ---- lifecycle -------------------------------------------------------------------------- */
/* Replaces EB:934-1040 + all cudaMalloc/cudaMemset of EB:1115-1357 (done lazily once graph,
 * features and labels are known). */
int gatx_create(gatx_ctx** out, const gatx_config* cfg);
void gatx_destroy(gatx_ctx* ctx);
const char* gatx_last_error(const gatx_ctx* ctx);
const char* gatx_version(void);
/* Number of usable devices (sm_100 class, ordinals 0..n-1); -1 when the CUDA runtime cannot be initialised.  A
 * multi-rank host checks it BEFORE creating contexts: a rank that cannot be created would leave its peers waiting in
 * the communicator (replaces nothing in the reference, which is single-GPU and never calls cudaSetDevice). */
int gatx_device_count(void);

/* ---- data ------------------------------------------------------------------------------- */
/* Replaces EB:1158-1192 (H2D of CSR + csr_to_coo_kernel) and adds the transposed graph and
 * the destination-row partition.  row_ptr [N+1], col_idx [E] describe the GLOBAL graph on every
 * rank; a rank keeps rows [bounds[rank], bounds[rank+1]). */
int gatx_set_graph_csr(gatx_ctx* ctx, int32_t num_nodes, int64_t num_edges,
                       const int32_t* row_ptr, const int32_t* col_idx);
/* Replaces EB:1151-1155.  X is the GLOBAL row-major [N][in_dim] matrix.  With world > 1 and an initialised
 * communicator the call is COLLECTIVE: every rank copies only its own rows host->device and the row blocks are
 * all-gathered over NVLink (every rank keeps all input rows, so layer 0 needs no exchange step). */
int gatx_set_features(gatx_ctx* ctx, const float* X, int32_t in_dim);
/* Replaces EB:1169-1172 + EB:1106-1107 (num_classes <= 0: derived as max(label)+1).  Labels outside [0, num_classes)
 * are refused with GATX_ERR_INVALID (the reference would index its class arrays out of bounds, EB:524, EB:572). */
int gatx_set_labels(gatx_ctx* ctx, const int32_t* labels, int32_t num_classes);
/* Extension (README.md:134 announces train/val/test splits "later"; the reference trains and scores on every
 * node, EB:514-550).  mask is a GLOBAL uint8 [N]: nodes with mask[n] == 0 contribute nothing to the loss, the
 * accuracy (both averaged over the counted nodes) or the gradients (dz = 0).  NULL restores the reference
 * behaviour.  A new graph drops the mask. */
int gatx_set_train_mask(gatx_ctx* ctx, const uint8_t* mask);
/* max row length (EB:89-99) and number of classes (EB:1106-1107) as the reference prints them */
int gatx_graph_info(gatx_ctx* ctx, int32_t* max_degree, int32_t* num_classes,
                    int32_t* row_begin, int32_t* row_end);
/* Destination-row partition bounds [world+1] for a global row_ptr (host-side helper). */
int gatx_partition_rows(int32_t num_nodes, const int32_t* row_ptr, int32_t world,
                        int32_t* bounds);

/* Blocks of the pipelined exchange (host-side helper, new work: the reference is single-GPU): every rank's own rows
 * [bounds[p], bounds[p+1]) cut into num_blocks edge-balanced blocks, out[p][k] = first GLOBAL row of block k of rank p,
 * out[p][num_blocks] = bounds[p+1].  The block table is part of the exchange protocol (every rank derives the same one
 * from the global row_ptr).  out holds world * (num_blocks + 1) entries. */
int gatx_row_blocks(int32_t num_nodes, const int32_t* row_ptr, int32_t world, int32_t num_blocks, int32_t* out);

/* ---- parameters ------------------------------------------------------------------------- */
/* Replaces setup_states_kernel + xavier_init_kernel_curand (EB:181-248, launch EB:1300-1323):
 * same distributions (limits EB:208, EB:236), counter-based and reproducible from `seed`. */
int gatx_init_params(gatx_ctx* ctx, uint64_t seed);
int gatx_set_params(gatx_ctx* ctx, int32_t layer, const float* W, const float* a);
int gatx_set_wo(gatx_ctx* ctx, const float* Wo);

/* Extension: the two LeakyReLU slopes, both fixed at 0.01 in the reference -- attn_slope inside the
 * attention score a . LReLU(W_l x_j + W_r x_i) (EB:1143; GATv2's negative_slope, 0.2 in the paper) and act_slope of
 * the layer activation (EB:1428).  0 <= slope < 1 (0 = ReLU).  Not calling it keeps the reference behaviour. */
int gatx_set_slopes(gatx_ctx* ctx, float attn_slope, float act_slope);
/* Extension (the reference has no dropout): inverted dropout with probability p on the INPUT of every
 * layer (features, then each hidden representation) in training forwards -- gatx_forward / gatx_train_epoch; never in
 * gatx_evaluate.  Element (global node n, column c) of layer l in the k-th training forward since this call is kept
 * iff word c % 4 of Philox4x32-10(counter {c / 4, n, l, k}, key seed) >= floor(p * 2^32), and scaled by 1 / (1 - p):
 * reproducible, identical on every rank of a partitioned run, re-generated (not stored) by the backward pass.
 * p = 0 switches it off.  0 <= p < 1. */
int gatx_set_dropout(gatx_ctx* ctx, float p, uint64_t seed);
/* Extension (the reference has no dropout): dropout with probability p on the ATTENTION COEFFICIENTS in
 * training forwards: h_i = sum_j alpha_ij d_ij W_l x_j with d_ij = keep / (1 - p) per (edge, head); the softmax itself
 * (EB:326-384) keeps every edge.  (edge e = GLOBAL CSR position, head h) of layer l in the k-th training forward since
 * this call is kept iff word h % 4 of Philox4x32-10(counter {e, h / 4, 2^31 | l, k}, key seed) >= floor(p * 2^32):
 * identical on every rank of a partitioned run, re-generated (not stored) by the backward pass.  Never applied in
 * gatx_evaluate.  p = 0 switches it off.  0 <= p < 1. */
int gatx_set_attn_dropout(gatx_ctx* ctx, float p, uint64_t seed);
/* Extension (the reference has no bias): a learnable bias b_l [H*D] per layer added to the aggregate
 * before the activation, h_i = sum_j alpha_ij W_l x_j + b_l (rows without in-edges give LReLU(b_l)).  The biases are
 * appended to the flat parameter / state buffer after W_o ([.. | W_o | b_0 .. b_{L-1}]), start at zero in
 * gatx_init_params, form a clip group of their own and are updated by the same Adam / SGD rule.  Switching it changes
 * the parameter layout: call it before the parameters are set (it drops them otherwise). */
int gatx_set_bias(gatx_ctx* ctx, int32_t on);
int gatx_set_bias_values(gatx_ctx* ctx, int32_t layer, const float* b /* [H*D] */);

/* ---- the epoch (EB:1370-1642) ----------------------------------------------------------- */
/* EB:1375-1452: per layer projection + score + segmented softmax + aggregation + activation,
 * then classifier + softmax. */
int gatx_forward(gatx_ctx* ctx);
/* EB:1455-1460 (compute_loss_accuracy_kernel + the two thrust reductions); synchronises. */
int gatx_loss_acc(gatx_ctx* ctx, float* avg_loss, float* accuracy);
/* EB:1463-1557. */
int gatx_backward(gatx_ctx* ctx);
/* EB:1560-1637: optional clip (3 groups), Adam (t = 1-based epoch) or SGD, zero gradients. */
int gatx_step(gatx_ctx* ctx, int32_t t);
/* forward + loss/accuracy + backward + step in one call, asynchronous until the scalars are
 * read; avg_loss/accuracy may be NULL (then no host sync at all). */
int gatx_train_epoch(gatx_ctx* ctx, int32_t t, float* avg_loss, float* accuracy);
/* Evaluation-only forward (extension): all layers + classifier + softmax, loss / accuracy over the
 * nodes of `mask` (GLOBAL uint8 [N], NULL = every node); no attention-coefficient storage, no output gradients,
 * parameters and optimizer state untouched.  It overwrites the activations, so gatx_backward needs a new
 * gatx_forward afterwards (GATX_ERR_INVALID otherwise).  Synchronises. */
int gatx_evaluate(gatx_ctx* ctx, const uint8_t* mask, float* avg_loss, float* accuracy);
int gatx_sync(gatx_ctx* ctx);
/* CUDA-graph replay of gatx_train_epoch (replaces the ~20 cudaDeviceSynchronize-separated launches per layer of
 * EB:1375-1557 with ONE graph launch per epoch).  Small graphs run tens of microsecond-sized kernels per epoch and
 * are bound by launch latency; forward + backward are captured once and replayed, the optimizer (t-dependent) and
 * the loss read-back follow as ordinary launches.  mode: -1 auto (default: single rank, <= 8 Mi edges, timing off;
 * env GATX_CUDA_GRAPH=0/1 overrides), 0 off, 1 on.  Results are bit-identical to the eager launches.  A new graph,
 * feature matrix, label set or train mask re-captures automatically. */
int gatx_set_cuda_graph(gatx_ctx* ctx, int32_t mode);
int gatx_cuda_graph_active(const gatx_ctx* ctx); /* 1 when the last gatx_train_epoch was a graph replay */

/* ---- introspection (parity tests, checkpoints) ------------------------------------------- */
int64_t gatx_tensor_size(gatx_ctx* ctx, int32_t which, int32_t layer); /* elements, <0 on error */
int gatx_get_tensor(gatx_ctx* ctx, int32_t which, int32_t layer, void* dst, size_t bytes);
/* Device milliseconds of the phases of the last gatx_train_epoch / forward / backward call:
 * out[0]=projection GEMMs, [1]=edge forward, [2]=classifier+loss, [3]=edge backward,
 * [4]=gradient GEMMs, [5]=optimizer, [6]=exchange (NCCL path: the collectives; peer-memory path: the time the compute
 * stream WAITS for the exchange stream, i.e. the exposed part), [7]=whole epoch.  Needs
 * gatx_enable_timing(ctx, 1) (adds event records, no host syncs). */
int gatx_enable_timing(gatx_ctx* ctx, int32_t on);
int gatx_get_timing(gatx_ctx* ctx, float* out_ms, int32_t n);
/* Device milliseconds of the three main streaming edge kernels of `layer` in the last epoch (timing enabled):
 * out3[0] = fused edge forward, [1] = backward pass 1 (destination-major), [2] = backward pass 2 (source-major);
 * zeros when the layer uses the narrow-row kernels. */
int gatx_get_edge_kernel_ms(gatx_ctx* ctx, int32_t layer, float* out3);
/* CUDA-event stopwatch on the context's launching stream (bench.py times its K steps with it). */
int gatx_timer_start(gatx_ctx* ctx);
int gatx_timer_stop(gatx_ctx* ctx, float* elapsed_ms); /* records, synchronises, returns ms */
/* Kernels launched by this context since creation (the bench's gpu_launches claim). */
int64_t gatx_launch_count(const gatx_ctx* ctx);
/* Algorithmic bytes of the fused edge forward / backward of one layer (SURVEY 8d formulas). */
int gatx_edge_bytes(gatx_ctx* ctx, int32_t layer, double* fwd_bytes, double* bwd_bytes);

/* ---- checkpoint (the reference only has dead dump/load helpers, NB:39-68) ------ */
/* Flat state = [parameters | Adam m | Adam v], each in the order W_0..W_{L-1} | a_0..a_{L-1} | W_o. */
int64_t gatx_state_size(gatx_ctx* ctx); /* floats, <0 on error */
int gatx_get_state(gatx_ctx* ctx, float* dst, size_t bytes);
int gatx_set_state(gatx_ctx* ctx, const float* src, size_t bytes);

/* ---- op-level entry point (per-kernel parity tests, ncu) ---------------------------------- */
/* Dense contraction on host buffers with the engine's GEMM kernels (mode = GATX_GEMM_*):
 *   form 0: C[M][N] = A[M][K] B[N][K]^T      (projection W x of EB:303-316, input gradient EB:859-869)
 *   form 1: C[M][N] = A[K][M]^T B[K][N]      (weight gradient of EB:771-782, contraction over nodes) */
int gatx_op_gemm(int32_t mode, int32_t form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                 int64_t ldc, int32_t M, int32_t N, int64_t K);

/* Fused edge forward of ONE layer on host buffers (EB:279-459: gatv2_edge_score_kernel, compute_max_sum_attn_score,
 * compute_attn_coeff, aggregate_kernel, postActivationLayerOutput) through the kernel family the epoch would pick for
 * (heads, outdim).  Pl, Pr [N][H*D] are the projected source / destination features, a [H*D]; any output may be NULL.
 * score, alpha [E][H] edge-major; hpre, Hout [N][H*D] (heads concatenated: hidden-layer semantics, EB:450-457). */
int gatx_op_edge_fwd(int32_t num_nodes, int64_t num_edges, const int32_t* row_ptr, const int32_t* col_idx, int32_t heads,
                     int32_t outdim, const float* Pl, const float* Pr, const float* a, float* score, float* alpha,
                     float* hpre, float* Hout);
/* Fused edge backward of ONE layer (EB:612-798 per-edge parts + EB:879-893): runs the forward above, then prep + pass 1 +
 * pass 2 with gHout [N][H*D] = dL/dHout.  Outputs (any may be NULL): g_pre [N][F] = gHout * LReLU'(h), gPl / gPr [N][F]
 * gradients w.r.t. the projected features, ga [F], ge [E][H] gradient w.r.t. the attention logits. */
int gatx_op_edge_bwd(int32_t num_nodes, int64_t num_edges, const int32_t* row_ptr, const int32_t* col_idx, int32_t heads,
                     int32_t outdim, const float* Pl, const float* Pr, const float* a, const float* gHout, float* g_pre,
                     float* gPl, float* gPr, float* ga, float* ge);
/* Softmax + cross-entropy + first-max argmax + dz = y - onehot on logits z [N][C] (EB:132-141, 514-550, 566-572); mask
 * (uint8 [N], NULL = all nodes) as in gatx_set_train_mask.  loss_sum = sum of -log(max(p, 1e-12)) over counted nodes. */
int gatx_op_softmax_ce(int32_t num_nodes, int32_t num_classes, const float* z, const int32_t* labels, const uint8_t* mask,
                       float* y, float* dz, int32_t* pred, double* loss_sum, int64_t* correct);
/* One optimizer step on a flat parameter vector (EB:250-278 clip per group at 5.0, EB:896-923 Adam / SGD, EB:1631-1633
 * gradient reset): group_end3 = end offsets of the three clip groups {all W}, {all a}, {W_o} (group_end3[2] = n); t is the
 * 1-based epoch.  params / grads / adam_m / adam_v are updated in place (adam_* may be NULL for SGD). */
int gatx_op_optimizer(int64_t n, const int64_t* group_end3, int32_t optimizer, int32_t clip, float lr, float beta1,
                      float beta2, int32_t t, float* params, float* grads, float* adam_m, float* adam_v);

/* ---- multi-GPU (one context per rank, NCCL over NVLink) --------------------------------- */
/* 128-byte NCCL unique id made on rank 0, distributed by the caller (torchrun store, MPI, file) */
int gatx_comm_unique_id(void* out128);
int gatx_comm_init(gatx_ctx* ctx, const void* id128);

/* ---- halo exchange through NVLink peer memory (optional) --------------------------------- */
/* With the destination-row partition a rank only needs the projected features of the sources its own edges gather
 * (the halo), and only the ranks that reference a source hold a partial gradient for it.  After these two calls the
 * per-layer exchange steps stop being NCCL collectives over whole [N][F] matrices: the owner of a row stores it
 * straight into the P_l buffers of the peers that reference it, and reads their partial gP_l rows back in a fixed
 * rank order (deterministic), all over NVLink load/store.  Call order: gatx_comm_init, graph / features / labels,
 * gatx_peer_export on every rank, exchange the GATX_PEER_INFO_BYTES blobs (any transport: files, MPI, torchrun
 * store), gatx_peer_import with all blobs in rank order.  Ranks may be processes (CUDA IPC handles) or threads of
 * one process (peer access on raw pointers).  Without these calls, or with GATX_NO_P2P set, the NCCL collectives run.
 * Changing the graph, features or labels invalidates the exchange (export / import again). */
#define GATX_PEER_INFO_BYTES 2048
int gatx_peer_export(gatx_ctx* ctx, void* out, size_t bytes);
int gatx_peer_import(gatx_ctx* ctx, const void* all_blobs, size_t bytes);
/* Back to the NCCL collectives on this rank.  The choice of exchange path must be the same on EVERY rank (different
 * sequences of barriers / collectives deadlock): when gatx_peer_export or gatx_peer_import fails on any rank, the host
 * calls this on all of them. */
int gatx_peer_disable(gatx_ctx* ctx);
/* Rows this rank pushes per layer (sum over own rows of the number of other ranks referencing them); an all-gather
 * would move (world - 1) * own rows.  Integer, identical to oracle/orc_halo_rows. */
int64_t gatx_halo_rows(const gatx_ctx* ctx);
int gatx_halo_active(const gatx_ctx* ctx); /* 1 when the peer-memory path is in use */
/* NVLink traffic of the exchange in the last epoch (timing enabled, peer-memory path): out4[0] = bytes this rank sent in
 * the forward exchange (projected rows into its peers' P_l buffers), [1] = milliseconds its busiest DMA stream (or its
 * push kernels, GATX_HALO_MODE=sm) spent on them, [2] / [3] = the same for the backward exchange (partial gP_l rows into
 * the owners' staging buffers).  The transfers run underneath the edge passes, so these are busy times, not exposed
 * times (phase 6 of gatx_get_timing is the exposed wait). */
int gatx_halo_stats(gatx_ctx* ctx, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* GATX_H_ */

// C ABI (include/gatx.h) and epoch orchestration of the B200-native GATv2 engine.
// Replaces the host side of the reference's main(): allocation EB:1115-1357 and the epoch loop
// EB:1370-1642 (one stream, no per-kernel cudaDeviceSynchronize, no host round trips except the
// two loss/accuracy scalars).
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gatx.h"
#include "common.cuh"
#include "gemm_tc.cuh"

using namespace gatx;

namespace {

// ---- NCCL through dlopen: single-GPU use needs no NCCL at all ------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return false;
#define L(sym) *(void**)(&sym) = dlsym(h, "nccl" #sym)
    L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(AllReduce); L(Broadcast); L(Reduce);
    L(GroupStart); L(GroupEnd); L(GetErrorString);
#undef L
    return GetUniqueId && CommInitRank && AllReduce && Broadcast && Reduce && GroupStart && GroupEnd;
  }
};
NcclApi g_nccl;

enum Phase { PH_GEMM_FWD = 0, PH_EDGE_FWD, PH_HEAD, PH_EDGE_BWD, PH_GEMM_BWD, PH_OPT, PH_COMM, PH_EPOCH, PH_COUNT };

struct Layer {
  int H = 0, D = 0, F = 0, I = 0, Fout = 0, ldx = 0, ldk = 0, recw = 0;
  bool vec = true;
  int64_t w_off = 0, a_off = 0, b_off = -1;  // b_off: bias of this layer in the flat buffer (gatx_set_bias), -1 = none
  float *Wcat = nullptr, *WcatT = nullptr;
  float* Xd = nullptr;  // dropped input of the last training forward (gatx_set_dropout), allocated on first use
  float *Pl = nullptr, *Pr = nullptr, *Hfull = nullptr, *Hout = nullptr, *hpre = nullptr;
  float *score = nullptr, *mx = nullptr, *sinv = nullptr, *gH = nullptr, *gHout = nullptr;
  float *galpha_dbg = nullptr, *gPl_dbg = nullptr, *gPr_dbg = nullptr, *alpha_dbg = nullptr, *ge_dbg = nullptr;
  cudaEvent_t kev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool kev_fwd = false, kev_bwd = false;  // recorded since timing was enabled
  bool gh_prepared = false;  // the input-gradient GEMM above already wrote g_pre into gH and the segment sums into cdot
};

}  // namespace

struct gatx_ctx {
  // config
  int L = 0, optimizer = 0, clip = 0, device = 0, gemm_mode = 0, keep_debug = 0, rank = 0, world = 1;
  float lr = 1e-4f, b1 = 0.9f, b2 = 0.999f;
  Slopes slopes{kSlope, kSlope};  // gatx_set_slopes; the reference's 0.01 / 0.01 by default
  // dropout on every layer's input in training forwards (gatx_set_dropout; the reference has none)
  float p_drop = 0.f;
  uint64_t drop_seed = 0;
  int64_t drop_step = 0;     // training forwards since gatx_set_dropout
  bool fwd_dropped = false;  // the last training forward used dropout (the backward must use the same inputs / mask)
  // attention-coefficient dropout (gatx_set_attn_dropout; the reference has none): the scale [E][H] of one layer is
  // generated into a scratch buffer right before the kernels that read it (forward, then again in the backward)
  float p_adrop = 0.f;
  uint64_t adrop_seed = 0;
  int64_t adrop_step = 0;
  bool fwd_adropped = false;
  float* ascale = nullptr;   // [E][Hmax]
  int64_t ascale_key = -1;   // (step, layer) the buffer currently holds
  int64_t edge0 = 0;         // global CSR position of this rank's first edge (the dropout draw is a function of it)
  // per-layer bias on the aggregate (gatx_set_bias): extra parameters [b_0 .. b_{L-1}] appended after W_o
  int use_bias = 0;
  int64_t bias_begin = 0, bias_end = 0;
  float* colsum_partials = nullptr;
  std::vector<int> heads, outdims;
  std::string err;
  cudaStream_t st = nullptr;
  int64_t launches = 0;
  // graph
  int N = 0;          // global nodes
  int64_t Eg = 0;     // global edges
  int r0 = 0, r1 = 0; // owned destination rows
  int n_rows = 0;
  int64_t E = 0;      // local edges
  int max_degree = 0;
  std::vector<int> bounds;
  int *row_ptr = nullptr, *col_idx = nullptr, *coo_src = nullptr, *coo_dst = nullptr, *in_deg = nullptr;
  int *csc_ptr = nullptr, *csc_dst = nullptr, *csc_eid = nullptr, *heavy_rows = nullptr, *heavy_srcs = nullptr;
  int n_heavy_rows = 0, n_heavy_srcs = 0;
  int chunk_T = 256, n_chunks = 0;
  int *chunk_row = nullptr, *chunk_src = nullptr;
  // multi-GPU pipeline: the own destination rows cut into edge-balanced blocks, each a graph of its own for the
  // streaming kernels (rebased row_ptr, own chunk table); one block = the whole range when pipelining is off
  struct RowBlock {
    int r0 = 0, n_rows = 0;
    int64_t e0 = 0, E = 0, halo_rows = 0;  // halo_rows: (own row, other rank that gathers it) pairs of the block
    int n_chunks = 0;
    int *row_ptr = nullptr, *chunk_row = nullptr;  // views into blk_row_ptr / blk_chunk_row
  };
  std::vector<RowBlock> blocks;
  int *blk_row_ptr = nullptr, *blk_chunk_row = nullptr;
  // L2 residency hints: index copies whose top bits mark the most frequently gathered nodes (nullptr = off)
  int *col_idx_hot = nullptr, *csc_dst_hot = nullptr;
  int hot_wide_F = 0;
  bool use_stream = true;
  bool have_graph = false, have_feat = false, have_labels = false, have_bufs = false, have_params = false;
  // data
  int I0 = 0, ld0 = 0, C = 0;
  float* X0 = nullptr;
  int* labels = nullptr;
  // optional node masks (extension): local rows, 1 = node counts; nullptr = every node (the reference)
  unsigned char *train_mask = nullptr, *eval_mask = nullptr;
  int64_t train_count = 0, eval_count = 0;  // GLOBAL number of counted nodes
  int64_t last_count = 0;                   // denominator of the last forward's loss / accuracy
  bool fwd_valid = false;                   // a training forward is the most recent forward
  bool eval_mode = false;                   // forward only: eval_mask in the loss, no output gradients
  // params: flat [W_0..W_{L-1} | a_0..a_{L-1} | W_o]
  int64_t n_params = 0, wo_off = 0;
  OptimGroups grp{};
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  std::vector<Layer> layers;
  // scratch
  float *gPl = nullptr, *gPr = nullptr, *ga_partials = nullptr, *splitk_ws = nullptr, *norm_partials = nullptr;
  float* x3_ws = nullptr;  // operand splits of the 3xTF32 GEMM mode (grown on demand)
  size_t x3_floats = 0;
  float* gPr2 = nullptr;  // second gP_r scratch: the pipelined backward writes layer l - 1's while layer l's is still read
  uint32_t* rec = nullptr;
  float *part = nullptr, *cdot = nullptr;
  size_t splitk_ws_bytes = 0;
  float *y = nullptr, *dz = nullptr, *z_dbg = nullptr, *WoT = nullptr;
  int ldc = 0;  // row pitch of y / dz / z: classes rounded up to 4 floats
  int* pred = nullptr;
  double *loss_partials = nullptr, *loss_sum = nullptr;  // loss_sum[0] = sum loss, [1] = correct (as double)
  int* correct_partials = nullptr;
  long long* correct = nullptr;
  double* red2 = nullptr;  // [2] all-reduce staging
  // comm
  ncclComm_t comm = nullptr;
  // NVLink peer-memory halo exchange (gatx_peer_export / gatx_peer_import); NCCL collectives when not set up
  uint16_t* ref_mask = nullptr;  // [n_rows] bit p: rank p's edge slice references this (own) source row
  int64_t halo_rows = 0;         // sum over own rows of the number of OTHER ranks referencing them
  bool peers_ready = false;
  std::vector<PeerPtrs> peer_Pl;  // per layer
  // backward exchange: every rank scatters its partial gP_l rows into slot `sender` of the owner's staging buffer
  float* stage = nullptr;          // [world][n_rows][Fmax]
  PeerPtrs peer_stage{};
  unsigned char* my_ref = nullptr; // [N] 1 = one of this rank's edges gathers the source (its partial row is non-zero)
  std::vector<int> all_blk;        // [world][K + 1] global row bounds of every rank's blocks
  std::vector<int64_t> scatter_rows;  // [K] rows this rank sends for block b (all owners)
  std::vector<void*> ipc_opened;
  // exchange stream + flag barriers in peer memory (halo_p2p.cu)
  cudaStream_t st_comm = nullptr;
  // Transport of the exchange.  Default: ld / st kernels -- forward the owner pushes its rows into the peers that gather
  // them, backward the owner pulls and sums the peers' partial rows.  GATX_HALO_MODE=bulk: bulk-copy (TMA) kernels in both
  // directions (forward push, backward scatter into the owners' staging buffers + a local ordered sum).
  bool halo_bulk = false;
  PeerPtrs peer_gPl{};
  uint32_t* halo_flags = nullptr;  // [kMaxPeers] slot p: the last barrier rank p has reached
  PeerFlags peer_flags{};
  uint32_t barrier_seq = 0;
  std::vector<cudaEvent_t> ev_pool;  // cross-stream ordering events, reused every epoch
  size_t ev_used = 0;
  struct CommSpan { int dir; double bytes; cudaEvent_t a, b; int lane = 0; };  // dir 0 = forward, 1 = backward; lane = DMA stream
  std::vector<CommSpan> comm_spans;
  size_t comm_spans_used = 0;
  // timing
  cudaEvent_t sw_a = nullptr, sw_b = nullptr;
  bool timing = false;
  struct Span { int phase; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t spans_used = 0;
  float phase_ms[PH_COUNT] = {0};
  // CUDA-graph replay of forward + backward (launch-bound shapes: tens of microsecond-sized kernels per epoch)
  int graph_mode = -1;            // -1 auto (single rank, at most kGraphAutoMaxEdges edges), 0 off, 1 on
  uint64_t gen = 1;               // bumped by everything that changes device pointers or the launch sequence
  cudaGraphExec_t epoch_exec = nullptr;
  uint64_t epoch_exec_gen = 0;
  int64_t epoch_exec_launches = 0, epoch_exec_count = 0;
  bool epoch_replayed = false;    // the last gatx_train_epoch ran as a graph launch
};

namespace {

int fail(gatx_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? GATX_ERR_OOM : GATX_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
  } while (0)
#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess)                                                                         \
      return fail(ctx, GATX_ERR_NCCL, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
  } while (0)
#define LAUNCHED(expr)                                                                             \
  do {                                                                                             \
    int n_ = (expr);                                                                               \
    if (n_ < 0) return fail(ctx, GATX_ERR_UNSUPPORTED, "%s failed", #expr);                        \
    ctx->launches += n_;                                                                           \
  } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  return cudaMalloc((void**)p, sizeof(T) * (n ? n : 1));
}
template <typename T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

struct PhaseTimer {
  gatx_ctx* c;
  size_t idx = (size_t)-1;
  PhaseTimer(gatx_ctx* ctx, int phase) : c(ctx) {
    if (!c->timing) return;
    if (c->spans_used == c->spans.size()) {
      gatx_ctx::Span s{phase, nullptr, nullptr};
      cudaEventCreate(&s.a);
      cudaEventCreate(&s.b);
      c->spans.push_back(s);
    }
    idx = c->spans_used++;
    c->spans[idx].phase = phase;
    cudaEventRecord(c->spans[idx].a, c->st);
  }
  ~PhaseTimer() {
    if (idx != (size_t)-1) cudaEventRecord(c->spans[idx].b, c->st);
  }
};

void free_graph(gatx_ctx* c) {
  dfree(c->row_ptr); dfree(c->col_idx); dfree(c->coo_src); dfree(c->coo_dst); dfree(c->in_deg);
  dfree(c->csc_ptr); dfree(c->csc_dst); dfree(c->csc_eid); dfree(c->heavy_rows); dfree(c->heavy_srcs);
  dfree(c->chunk_row); dfree(c->chunk_src); dfree(c->ref_mask); dfree(c->col_idx_hot); dfree(c->csc_dst_hot);
  dfree(c->blk_row_ptr); dfree(c->blk_chunk_row); dfree(c->my_ref);
  c->blocks.clear();
  c->all_blk.clear();
  c->scatter_rows.clear();
  c->have_graph = false;
  ++c->gen;
}
void free_bufs(gatx_ctx* c) {
  ++c->gen;
  for (void* q : c->ipc_opened) cudaIpcCloseMemHandle(q);
  c->ipc_opened.clear();
  c->peers_ready = false;
  dfree(c->halo_flags);
  dfree(c->gPr2);
  dfree(c->stage);
  for (auto& l : c->layers) {
    dfree(l.Wcat); dfree(l.WcatT); dfree(l.Pl); dfree(l.Pr); dfree(l.Xd);
    if (l.Hout != l.Hfull) dfree(l.Hout);
    l.Hout = nullptr;
    dfree(l.Hfull); dfree(l.hpre); dfree(l.score); dfree(l.mx); dfree(l.sinv); dfree(l.gH); dfree(l.gHout);
    dfree(l.galpha_dbg); dfree(l.gPl_dbg); dfree(l.gPr_dbg); dfree(l.alpha_dbg); dfree(l.ge_dbg);
    for (auto& e : l.kev) {
      if (e) cudaEventDestroy(e);
      e = nullptr;
    }
    l.kev_fwd = l.kev_bwd = false;
  }
  dfree(c->params); dfree(c->grads); dfree(c->adam_m); dfree(c->adam_v);
  dfree(c->x3_ws);
  c->x3_floats = 0;
  dfree(c->gPl); dfree(c->gPr); dfree(c->ga_partials); dfree(c->splitk_ws); dfree(c->norm_partials); dfree(c->colsum_partials);
  dfree(c->ascale);
  c->ascale_key = -1;
  dfree(c->rec); dfree(c->part); dfree(c->cdot); dfree(c->y); dfree(c->dz); dfree(c->z_dbg); dfree(c->WoT); dfree(c->pred);
  dfree(c->loss_partials); dfree(c->loss_sum); dfree(c->correct_partials); dfree(c->correct); dfree(c->red2);
  c->have_bufs = false;
  c->have_params = false;
}

// All device state that depends on (graph, in_dim, classes): EB:1196-1357.
int ensure_buffers(gatx_ctx* ctx) {
  if (ctx->have_bufs) return GATX_OK;
  if (!ctx->have_graph || !ctx->have_feat || !ctx->have_labels)
    return fail(ctx, GATX_ERR_INVALID, "graph, features and labels must be set first");
  const int L = ctx->L, N = ctx->N, nr = ctx->n_rows;
  const int64_t E = ctx->E;
  int64_t off = 0, Fmax = 0, recmax = 0;
  for (int l = 0; l < L; ++l) {
    Layer& ly = ctx->layers[l];
    ly.H = ctx->heads[l];
    ly.D = ctx->outdims[l];
    ly.F = ly.H * ly.D;
    ly.I = l == 0 ? ctx->I0 : ctx->layers[l - 1].Fout;  // EB:1115-1118
    ly.ldx = l == 0 ? ctx->ld0 : ctx->layers[l - 1].Fout;
    // EB semantics: hidden layers concatenate heads; the last layer averages them (EB:440-459)
    ly.Fout = (l == L - 1) ? ly.D : ly.F;
    ly.ldk = (ly.I + 3) / 4 * 4;
    ly.vec = edge_shape_supported(ly.H, ly.D);  // vectorised kernels; otherwise the generic scalar kernels
    ly.recw = ly.vec ? edge_rec_words(ly.H, ly.D) : edge_generic_rec_words(ly.H);
    if (!ly.vec && !edge_generic_supported(ly.H, ly.D))
      return fail(ctx, GATX_ERR_UNSUPPORTED,
                  "layer %d: heads=%d outdim=%d not covered by the edge kernels (need heads <= 32 and "
                  "heads*outdim <= 1024)", l, ly.H, ly.D);
    ly.w_off = off;
    off += (int64_t)ly.F * 2 * ly.I;
    if (ly.F > Fmax) Fmax = ly.F;
    if (ly.recw > recmax) recmax = ly.recw;
  }
  ctx->grp.begin[0] = 0;
  ctx->grp.end[0] = off;
  ctx->grp.begin[1] = off;
  for (int l = 0; l < L; ++l) {
    ctx->layers[l].a_off = off;
    off += ctx->layers[l].F;
  }
  ctx->grp.end[1] = off;
  ctx->grp.begin[2] = off;
  ctx->wo_off = off;
  const int DL = ctx->outdims[L - 1];
  off += (int64_t)ctx->C * DL;
  ctx->grp.end[2] = off;
  ctx->bias_begin = off;
  for (int l = 0; l < L; ++l) {
    ctx->layers[l].b_off = ctx->use_bias ? off : -1;
    if (ctx->use_bias) off += ctx->layers[l].F;
  }
  ctx->bias_end = off;
  ctx->n_params = off;
  CK(dalloc(&ctx->params, off));
  CK(dalloc(&ctx->grads, off));
  CK(dalloc(&ctx->adam_m, off));
  CK(dalloc(&ctx->adam_v, off));
  CK(cudaMemsetAsync(ctx->params, 0, sizeof(float) * off, ctx->st));
  CK(cudaMemsetAsync(ctx->grads, 0, sizeof(float) * off, ctx->st));   // EB:1262-1266
  CK(cudaMemsetAsync(ctx->adam_m, 0, sizeof(float) * off, ctx->st));  // EB:1275-1295
  CK(cudaMemsetAsync(ctx->adam_v, 0, sizeof(float) * off, ctx->st));
  for (int l = 0; l < L; ++l) {
    Layer& ly = ctx->layers[l];
    CK(dalloc(&ly.Wcat, (size_t)2 * ly.F * ly.ldk));
    CK(dalloc(&ly.WcatT, (size_t)ly.I * 2 * ly.F));
    CK(dalloc(&ly.Pl, (size_t)N * ly.F));
    CK(dalloc(&ly.Pr, (size_t)nr * ly.F));
    CK(dalloc(&ly.Hfull, (size_t)nr * ly.F));
    if (ly.Fout != ly.F) CK(dalloc(&ly.Hout, (size_t)nr * ly.Fout));
    else ly.Hout = ly.Hfull;
    CK(dalloc(&ly.score, (size_t)E * ly.H));
    CK(dalloc(&ly.mx, (size_t)nr * ly.H));
    CK(dalloc(&ly.sinv, (size_t)nr * ly.H));
    CK(dalloc(&ly.gH, (size_t)nr * ly.F));
    if (ly.Fout != ly.F) CK(dalloc(&ly.gHout, (size_t)nr * ly.Fout));
    if (ctx->keep_debug) {
      CK(dalloc(&ly.hpre, (size_t)nr * ly.F));
      CK(dalloc(&ly.galpha_dbg, (size_t)E * ly.H));
      CK(dalloc(&ly.alpha_dbg, (size_t)E * ly.H));
      CK(dalloc(&ly.ge_dbg, (size_t)E * ly.H));
      CK(dalloc(&ly.gPl_dbg, (size_t)N * ly.F));
      CK(dalloc(&ly.gPr_dbg, (size_t)nr * ly.F));
    }
  }
  CK(dalloc(&ctx->gPl, (size_t)N * Fmax));
  CK(dalloc(&ctx->gPr, (size_t)nr * Fmax));
  if (ctx->blocks.size() > 1) CK(dalloc(&ctx->gPr2, (size_t)nr * Fmax));
  if (ctx->world > 1) {
    CK(dalloc(&ctx->stage, (size_t)ctx->world * nr * Fmax));
    CK(dalloc(&ctx->halo_flags, (size_t)kMaxPeers));
    CK(cudaMemsetAsync(ctx->halo_flags, 0, sizeof(uint32_t) * kMaxPeers, ctx->st));
    ctx->barrier_seq = 0;
  }
  CK(dalloc(&ctx->rec, (size_t)E * recmax));
  {
    int64_t part_floats = 1, hmax = 1;
    for (int l = 0; l < L; ++l) {
      const int64_t pf = edge_stream_part_floats(ctx->layers[l].H, ctx->layers[l].D, ctx->n_chunks);
      if (pf > part_floats) part_floats = pf;
      if (ctx->layers[l].H > hmax) hmax = ctx->layers[l].H;
    }
    CK(dalloc(&ctx->part, (size_t)part_floats));
    CK(dalloc(&ctx->cdot, (size_t)nr * hmax));
  }
  CK(dalloc(&ctx->ga_partials, (size_t)(kNumSMs * 8 + ctx->n_heavy_rows + 1) * Fmax));
  ctx->splitk_ws_bytes = (size_t)256 << 20;
  CK(dalloc(&ctx->splitk_ws, ctx->splitk_ws_bytes / sizeof(float)));
  CK(dalloc(&ctx->norm_partials, (size_t)6 * kOptimBlocks));  // [0, 3): the reference's groups, [3, 6): the bias group
  if (ctx->use_bias) CK(dalloc(&ctx->colsum_partials, (size_t)kColsumBlocks * Fmax));
  ctx->ldc = (ctx->C + 3) / 4 * 4;
  CK(dalloc(&ctx->y, (size_t)nr * ctx->ldc));
  CK(dalloc(&ctx->dz, (size_t)nr * ctx->ldc));
  CK(dalloc(&ctx->z_dbg, (size_t)nr * ctx->ldc));
  CK(dalloc(&ctx->WoT, (size_t)DL * ctx->ldc));
  CK(dalloc(&ctx->pred, (size_t)nr));
  CK(dalloc(&ctx->loss_partials, (size_t)kHeadBlocks));
  CK(dalloc(&ctx->correct_partials, (size_t)kHeadBlocks));
  CK(dalloc(&ctx->loss_sum, 1));
  CK(dalloc(&ctx->correct, 1));
  CK(dalloc(&ctx->red2, 2));
  ctx->have_bufs = true;
  ++ctx->gen;
  return GATX_OK;
}

EdgeGraph edge_graph(const gatx_ctx* c) {
  EdgeGraph g{};
  g.n_rows = c->n_rows; g.row_ptr = c->row_ptr; g.col_idx = c->col_idx; g.n_src = c->N;
  g.csc_ptr = c->csc_ptr; g.csc_dst = c->csc_dst; g.csc_eid = c->csc_eid;
  g.heavy_rows = c->heavy_rows; g.n_heavy_rows = c->n_heavy_rows;
  g.heavy_srcs = c->heavy_srcs; g.n_heavy_srcs = c->n_heavy_srcs;
  g.E = c->E;
  g.chunk_T = c->chunk_T; g.n_chunks = c->n_chunks; g.chunk_row = c->chunk_row; g.chunk_src = c->chunk_src;
  g.col_idx_hot = c->col_idx_hot; g.csc_dst_hot = c->csc_dst_hot; g.hot_wide_F = c->hot_wide_F;
  g.kernel_events = nullptr;
  g.reserve_ctas = 0;
  g.slopes = c->slopes;
  g.bias = nullptr;
  g.ascale = nullptr;
  return g;
}

// P_l | P_r = X [W_l ; W_r]^T in one pass over X (Wcat is [2F][ldk], rows 0..F-1 = W_l) for `rows` rows of X.
// ---- 3xTF32: fp32-grade contractions on the tensor cores ------------------------------------------------------------
// x = hi + lo with hi = the TF32 part of x (what kind::tf32 reads: the low 13 mantissa bits dropped) and lo = x - hi
// (exact).  A B^T ~= A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T, all three accumulated in fp32 in the same TMEM tile by the
// existing kernels: the first two as ONE product over the concatenated contraction index ([A_hi | A_lo] . [B_hi | B_hi]),
// the third as the kernel's second operand pair.  The dropped lo * lo term is ~2^-22 of the product.
constexpr int kX3Fallback = 1000;
int x3_reserve(gatx_ctx* ctx, size_t floats, float** out) {
  if (ctx->x3_floats < floats) {
    CK(cudaStreamSynchronize(ctx->st));
    CK(dalloc(&ctx->x3_ws, floats));
    ctx->x3_floats = floats;
  }
  *out = ctx->x3_ws;
  return GATX_OK;
}
// C0 | C1 (split at n_split) (+)= A[M][K] B[N][K]^T
int gemm3x_tn(gatx_ctx* ctx, const float* A, int64_t lda, const float* B, int64_t ldb, int K, float* C0, float* C1,
              int n_split, int64_t ldc, int M, int N, bool accumulate) {
  const int Kp = (K + 3) / 4 * 4;
  float* ws;
  int rc = x3_reserve(ctx, (size_t)M * 2 * Kp + (size_t)N * 3 * Kp, &ws);
  if (rc) return rc;
  float* A2 = ws;                              // [M][2 Kp] = [A_hi | A_lo]
  float* B2 = A2 + (size_t)M * 2 * Kp;         // [N][2 Kp] = [B_hi | B_hi]
  float* Blo = B2 + (size_t)N * 2 * Kp;        // [N][Kp]
  LAUNCHED(launch_split_tf32(A, lda, M, K, Kp, A2, 2 * Kp, A2 + Kp, 2 * Kp, ctx->st));
  LAUNCHED(launch_split_tf32(B, ldb, N, K, Kp, B2, 2 * Kp, Blo, Kp, ctx->st));
  CK(cudaMemcpy2DAsync(B2 + Kp, sizeof(float) * 2 * Kp, B2, sizeof(float) * 2 * Kp, sizeof(float) * Kp, N,
                       cudaMemcpyDeviceToDevice, ctx->st));
  int n = launch_gemm_tc_tn2(A2, 2 * Kp, B2, 2 * Kp, 2 * Kp, A2, 2 * Kp, Blo, Kp, Kp, C0, C1, n_split, ldc, M, N,
                             accumulate, ctx->st);
  if (n < 0) return kX3Fallback;  // a shape the tensor-core kernel does not take: the caller runs the fp32 CUDA-core GEMM
  ctx->launches += n;
  return GATX_OK;
}
// C[M][N] += A[K][M]^T B[K][N] (contraction over nodes)
int gemm3x_atb(gatx_ctx* ctx, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
               int64_t K) {
  const int Mp = (M + 3) / 4 * 4, Np = (N + 3) / 4 * 4;
  float* ws;
  int rc = x3_reserve(ctx, (size_t)K * 2 * (Mp + Np), &ws);
  if (rc) return rc;
  float *Ahi = ws, *Alo = Ahi + (size_t)K * Mp, *Bhi = Alo + (size_t)K * Mp, *Blo = Bhi + (size_t)K * Np;
  LAUNCHED(launch_split_tf32(A, lda, K, M, Mp, Ahi, Mp, Alo, Mp, ctx->st));
  LAUNCHED(launch_split_tf32(B, ldb, K, N, Np, Bhi, Np, Blo, Np, ctx->st));
  const float* as[3] = {Alo, Ahi, Ahi};  // the small terms first
  const float* bs[3] = {Bhi, Blo, Bhi};
  for (int i = 0; i < 3; ++i) {
    int n = launch_gemm_tc_atb(as[i], Mp, bs[i], Np, C, ldc, M, N, K, ctx->splitk_ws, ctx->splitk_ws_bytes, ctx->st);
    if (n < 0) return i == 0 ? kX3Fallback : fail(ctx, GATX_ERR_UNSUPPORTED, "3xTF32: A^T B GEMM rejected M=%d N=%d", M, N);
    ctx->launches += n;
  }
  return GATX_OK;
}

int gemm_project(gatx_ctx* ctx, const float* X, int ldx, const Layer& ly, float* Pl_rows, float* Pr_rows, int rows) {
  if (rows <= 0) return GATX_OK;
  if (ctx->gemm_mode == GATX_GEMM_3XTF32_TC) {
    const int rc = gemm3x_tn(ctx, X, ldx, ly.Wcat, ly.ldk, ly.I, Pl_rows, Pr_rows, ly.F, ly.F, rows, 2 * ly.F, false);
    if (rc != kX3Fallback) return rc;
  }
  if (ctx->gemm_mode == GATX_GEMM_TF32_TC) {
    int n = launch_gemm_tc_tn2(X, ldx, ly.Wcat, ly.ldk, ly.I, nullptr, 0, nullptr, 0, 0, Pl_rows, Pr_rows, ly.F, ly.F,
                               rows, 2 * ly.F, false, ctx->st);
    if (n >= 0) {
      ctx->launches += n;
      return GATX_OK;
    }
  }
  LAUNCHED(launch_gemm_simt(X, ldx, 1, ly.Wcat, ly.ldk, 1, Pl_rows, ly.F, rows, ly.F, ly.I, false, nullptr, 0, ctx->st));
  LAUNCHED(launch_gemm_simt(X, ldx, 1, ly.Wcat + (int64_t)ly.F * ly.ldk, ly.ldk, 1, Pr_rows, ly.F, rows, ly.F, ly.I,
                            false, nullptr, 0, ctx->st));
  return GATX_OK;
}
// C[M][N] = A[M][K] B[N][K]^T with the configured arithmetic
int gemm_tn_any(gatx_ctx* ctx, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                int N, int K) {
  if (ctx->gemm_mode == GATX_GEMM_3XTF32_TC) {
    const int rc = gemm3x_tn(ctx, A, lda, B, ldb, K, C, C, N, ldc, M, N, false);
    if (rc != kX3Fallback) return rc;
  }
  if (ctx->gemm_mode == GATX_GEMM_TF32_TC) {
    int n = launch_gemm_tc_tn(A, lda, B, ldb, C, ldc, M, N, K, false, ctx->st);
    if (n >= 0) {
      ctx->launches += n;
      return GATX_OK;
    }
  }
  LAUNCHED(launch_gemm_simt(A, lda, 1, B, ldb, 1, C, ldc, M, N, K, false, nullptr, 0, ctx->st));
  return GATX_OK;
}
// gX[n][i] = sum_r gP_l[n][r] W_l[r][i] + gP_r[n][r] W_r[r][i]   (WcatT is [I][2F]) for `rows` rows
// `fuse` (optional): epilogue that turns the tile into the pre-activation gradient of the layer below and its segment sums
// (see GemmFuse); *fused tells the caller whether it ran (tensor-core path with 256-column tiles only).
int gemm_input_grad(gatx_ctx* ctx, const float* gPl_rows, const float* gPr_rows, const Layer& ly, float* gX, int ldg,
                    int rows, const GemmFuse* fuse = nullptr, bool* fused = nullptr) {
  if (fused) *fused = false;
  if (rows <= 0) return GATX_OK;
  if (ctx->gemm_mode == GATX_GEMM_TF32_TC && fuse && fuse->Hout) {
    int n = launch_gemm_tc_tn2(gPl_rows, ly.F, ly.WcatT, 2 * ly.F, ly.F, gPr_rows, ly.F, ly.WcatT + ly.F, 2 * ly.F, ly.F,
                               gX, gX, ly.I, ldg, rows, ly.I, false, ctx->st, fuse);
    if (n >= 0) {
      ctx->launches += n;
      if (fused) *fused = true;
      return GATX_OK;
    }
  }
  if (ctx->gemm_mode == GATX_GEMM_3XTF32_TC) {
    int rc = gemm3x_tn(ctx, gPl_rows, ly.F, ly.WcatT, 2 * ly.F, ly.F, gX, gX, ly.I, ldg, rows, ly.I, false);
    if (rc == GATX_OK) return gemm3x_tn(ctx, gPr_rows, ly.F, ly.WcatT + ly.F, 2 * ly.F, ly.F, gX, gX, ly.I, ldg, rows, ly.I, true);
    if (rc != kX3Fallback) return rc;
  }
  if (ctx->gemm_mode == GATX_GEMM_TF32_TC) {
    int n = launch_gemm_tc_tn2(gPl_rows, ly.F, ly.WcatT, 2 * ly.F, ly.F, gPr_rows, ly.F, ly.WcatT + ly.F, 2 * ly.F, ly.F,
                               gX, gX, ly.I, ldg, rows, ly.I, false, ctx->st);
    if (n >= 0) {
      ctx->launches += n;
      return GATX_OK;
    }
  }
  LAUNCHED(launch_gemm_simt(gPl_rows, ly.F, 1, ly.WcatT, 2 * ly.F, 1, gX, ldg, rows, ly.I, ly.F, false, nullptr, 0,
                            ctx->st));
  LAUNCHED(launch_gemm_simt(gPr_rows, ly.F, 1, ly.WcatT + ly.F, 2 * ly.F, 1, gX, ldg, rows, ly.I, ly.F, true, nullptr, 0,
                            ctx->st));
  return GATX_OK;
}
// C[M][N] (ldc) += A^T B with A [K][>=M] (lda), B [K][>=N] (ldb): contraction over the node dimension.
int gemm_nt_reduce(gatx_ctx* ctx, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int M, int N, int64_t K) {
  if (ctx->gemm_mode == GATX_GEMM_3XTF32_TC) {
    const int rc = gemm3x_atb(ctx, A, lda, B, ldb, C, ldc, M, N, K);
    if (rc != kX3Fallback) return rc;
  }
  if (ctx->gemm_mode == GATX_GEMM_TF32_TC) {
    int n = launch_gemm_tc_atb(A, lda, B, ldb, C, ldc, M, N, K, ctx->splitk_ws, ctx->splitk_ws_bytes, ctx->st);
    if (n >= 0) {
      ctx->launches += n;
      return GATX_OK;
    }
  }
  LAUNCHED(launch_gemm_simt(A, 1, lda, B, 1, ldb, C, ldc, M, N, K, true, ctx->splitk_ws, ctx->splitk_ws_bytes,
                            ctx->st));
  return GATX_OK;
}

// ---- multi-GPU exchange ------------------------------------------------------------------------------------------
// Peer-memory path (default): the exchange kernels run on a stream of their own (st_comm), one block of destination
// rows at a time, ordered against the compute stream by events and across ranks by flag barriers in peer memory, so
// that the NVLink traffic of block b overlaps the edge pass of block b + 1 (see do_forward / do_backward).
// NCCL path (GATX_NO_P2P, or no peer access): whole-matrix broadcasts / reduces on the compute stream.
bool halo_p2p(const gatx_ctx* ctx, int F) { return ctx->world > 1 && ctx->peers_ready && F % 4 == 0; }

cudaEvent_t next_event(gatx_ctx* c) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    c->ev_pool.push_back(e);
  }
  return c->ev_pool[c->ev_used++];
}
// `waiter` continues only after everything enqueued on `src` so far
void stream_after(gatx_ctx* c, cudaStream_t waiter, cudaStream_t src) {
  cudaEvent_t e = next_event(c);
  cudaEventRecord(e, src);
  cudaStreamWaitEvent(waiter, e, 0);
}
// the compute stream waits for the exchange stream; the wait is what the PH_COMM phase measures (EXPOSED exchange time)
void compute_waits_comm(gatx_ctx* ctx) {
  PhaseTimer t(ctx, PH_COMM);
  stream_after(ctx, ctx->st, ctx->st_comm);
}
int comm_barrier(gatx_ctx* ctx) {
  ++ctx->barrier_seq;
  LAUNCHED(launch_halo_barrier(ctx->peer_flags, ctx->rank, ctx->world, ctx->barrier_seq, ctx->st_comm));
  return GATX_OK;
}
// timing of the exchange kernels themselves (on st_comm): bytes over NVLink and busy time per direction
struct CommTimer {
  gatx_ctx* c;
  size_t idx = (size_t)-1;
  CommTimer(gatx_ctx* ctx, int dir, double bytes) : c(ctx) {
    if (!c->timing) return;
    if (c->comm_spans_used == c->comm_spans.size()) {
      gatx_ctx::CommSpan s{dir, 0.0, nullptr, nullptr};
      cudaEventCreate(&s.a);
      cudaEventCreate(&s.b);
      c->comm_spans.push_back(s);
    }
    idx = c->comm_spans_used++;
    c->comm_spans[idx].dir = dir;
    c->comm_spans[idx].bytes = bytes;
    c->comm_spans[idx].lane = 0;
    cudaEventRecord(c->comm_spans[idx].a, c->st_comm);
  }
  ~CommTimer() {
    if (idx != (size_t)-1) cudaEventRecord(c->comm_spans[idx].b, c->st_comm);
  }
};

// All-gather of row blocks with unequal counts: one NCCL broadcast per owner inside a group.
int nccl_allgather_rows(gatx_ctx* ctx, float* full, int64_t F) {
  NK(g_nccl.GroupStart());
  for (int r = 0; r < ctx->world; ++r) {
    float* p = full + (int64_t)ctx->bounds[r] * F;
    const size_t cnt = (size_t)(ctx->bounds[r + 1] - ctx->bounds[r]) * F;
    if (cnt) NK(g_nccl.Broadcast(p, p, cnt, ncclFloat, r, ctx->comm, ctx->st));
  }
  NK(g_nccl.GroupEnd());
  return GATX_OK;
}

// Forward exchange, NCCL path: every rank ends with all P_l rows.
int comm_allgather_rows(gatx_ctx* ctx, float* full, int F) {
  if (ctx->world == 1) return GATX_OK;
  if (!ctx->comm) return fail(ctx, GATX_ERR_INVALID, "world > 1 but gatx_comm_init was not called");
  PhaseTimer t(ctx, PH_COMM);
  return nccl_allgather_rows(ctx, full, F);
}

// Backward exchange, NCCL path: the owner of a source row ends with the sum of all ranks' partial gP_l rows.
int comm_reduce_rows(gatx_ctx* ctx, float* full, int F) {
  if (ctx->world == 1) return GATX_OK;
  if (!ctx->comm) return fail(ctx, GATX_ERR_INVALID, "world > 1 but gatx_comm_init was not called");
  PhaseTimer t(ctx, PH_COMM);
  NK(g_nccl.GroupStart());
  for (int r = 0; r < ctx->world; ++r) {
    float* p = full + (int64_t)ctx->bounds[r] * F;
    const size_t cnt = (size_t)(ctx->bounds[r + 1] - ctx->bounds[r]) * F;
    if (cnt) NK(g_nccl.Reduce(p, p, cnt, ncclFloat, ncclSum, r, ctx->comm, ctx->st));
  }
  NK(g_nccl.GroupEnd());
  return GATX_OK;
}

// ---- views of the own destination rows: everything (b < 0) or one block of the multi-GPU pipeline ---------------
struct RowView {
  int rb, nb;    // rows [rb, rb + nb) of this rank's own rows
  int64_t e0;    // first local edge of the view
  EdgeGraph g;   // the view as a graph of its own (rebased row_ptr, own chunks, col_idx offset to e0)
  bool whole;
};
RowView row_view(const gatx_ctx* ctx, int b) {
  RowView v{0, ctx->n_rows, 0, edge_graph(ctx), true};
  if (b < 0 || ctx->blocks.size() <= 1) return v;
  const gatx_ctx::RowBlock& B = ctx->blocks[b];
  v.rb = B.r0; v.nb = B.n_rows; v.e0 = B.e0; v.whole = false;
  v.g.n_rows = B.n_rows; v.g.row_ptr = B.row_ptr; v.g.col_idx = ctx->col_idx + B.e0; v.g.E = B.E;
  v.g.n_chunks = B.n_chunks; v.g.chunk_row = B.chunk_row;
  v.g.col_idx_hot = ctx->col_idx_hot ? ctx->col_idx_hot + B.e0 : nullptr;
  v.g.heavy_rows = nullptr; v.g.n_heavy_rows = 0;
  // SM transport: the exchange kernels of the neighbouring block run underneath this launch in their own CTA slots
  v.g.reserve_ctas = halo_cta_slots(ctx->world);
  return v;
}
// the streaming kernels take any view; the warp-per-row / generic kernels only the whole row range
bool blockable(const gatx_ctx* ctx, int l) {
  const Layer& ly = ctx->layers[l];
  return ctx->blocks.size() > 1 && ly.vec && ctx->use_stream && edge_stream_supported(ly.H, ly.D);
}
float* gpr_buf(const gatx_ctx* ctx, int l) { return (ctx->gPr2 && ((ctx->L - 1 - l) & 1)) ? ctx->gPr2 : ctx->gPr; }

// Attention-coefficient dropout: the scale of layer l for the current step, generated on the compute stream into the
// scratch buffer (whole layer at once; a block view reads its slice).  nullptr when the option is off.
int attn_scale(gatx_ctx* ctx, int l, bool active, const RowView& v, const float** out) {
  *out = nullptr;
  if (!active) return GATX_OK;
  const Layer& ly = ctx->layers[l];
  const int64_t key = ctx->adrop_step * 64 + l;
  if (!ctx->ascale) {
    int hmax = 1;
    for (const Layer& q : ctx->layers) hmax = q.H > hmax ? q.H : hmax;
    CK(dalloc(&ctx->ascale, (size_t)ctx->E * hmax));
    ctx->ascale_key = -1;
  }
  if (ctx->ascale_key != key) {
    LAUNCHED(launch_attn_dropout_scale(ctx->ascale, ctx->E, ly.H, ctx->edge0, ctx->p_adrop, ctx->adrop_seed, l,
                                       ctx->adrop_step, ctx->st));
    ctx->ascale_key = key;
  }
  *out = ctx->ascale + v.e0 * ly.H;
  return GATX_OK;
}

// fused edge forward of layer l over a view (EB:279-459)
int fwd_edge(gatx_ctx* ctx, int l, const RowView& v) {
  Layer& ly = ctx->layers[l];
  if (v.nb <= 0) return GATX_OK;
  EdgeGraph gl = v.g;
  gl.bias = ly.b_off >= 0 ? ctx->params + ly.b_off : nullptr;
  if (int rc = attn_scale(ctx, l, !ctx->eval_mode && ctx->p_adrop > 0.f, v, &gl.ascale)) return rc;
  if (ctx->timing && v.whole) {
    for (auto& e : ly.kev)
      if (!e) cudaEventCreate(&e);
    gl.kernel_events = ly.kev;
    ly.kev_fwd = true;
  }
  const int64_t ro = v.rb;
  float* Pr = ly.Pr + ro * ly.F;
  float* Hfull = ly.Hfull + ro * ly.F;
  float* hpre = ly.hpre ? ly.hpre + ro * ly.F : nullptr;
  float* score = ly.score + v.e0 * ly.H;
  float *mx = ly.mx + ro * ly.H, *sinv = ly.sinv + ro * ly.H;
  const float* a = ctx->params + ly.a_off;
  if (!ly.vec)
    LAUNCHED(launch_edge_forward_generic(gl, ly.H, ly.D, ly.Pl, Pr, a, Hfull, hpre, score, mx, sinv, ctx->st));
  else if (ctx->use_stream && edge_stream_supported(ly.H, ly.D))
    LAUNCHED(launch_edge_forward_stream(gl, ly.H, ly.D, ly.Pl, Pr, a, Hfull, hpre, score, mx, sinv, ctx->part, ctx->st));
  else
    LAUNCHED(launch_edge_forward(gl, ly.H, ly.D, ly.Pl, Pr, a, Hfull, hpre, score, mx, sinv, ctx->st));
  return GATX_OK;
}

// backward of layer l over a view: phases bit 0 = prep + destination-major pass 1 (over the view's rows),
// bit 1 = source-major pass 2 (always over all local edges; the view must be the whole range)
int bwd_edge(gatx_ctx* ctx, int l, const RowView& v, int phases) {
  Layer& ly = ctx->layers[l];
  EdgeGraph gl = v.g;
  gl.bias = ly.b_off >= 0 ? ctx->params + ly.b_off : nullptr;
  if (phases & 1)  // pass 1 re-generates the scale of the forward it differentiates (same step counter)
    if (int rc = attn_scale(ctx, l, ctx->fwd_adropped, v, &gl.ascale)) return rc;
  if (ctx->timing && v.whole) {
    for (auto& e : ly.kev)
      if (!e) cudaEventCreate(&e);
    gl.kernel_events = ly.kev;
    ly.kev_bwd = true;
  }
  const int64_t ro = v.rb;
  const float* a = ctx->params + ly.a_off;
  float* gPr = gpr_buf(ctx, l) + ro * ly.F;
  const float* Pr = ly.Pr + ro * ly.F;
  const float* Hfull = ly.Hfull + ro * ly.F;
  float* gH = ly.gH + ro * ly.F;
  const float* score = ly.score + v.e0 * ly.H;
  const float *mx = ly.mx + ro * ly.H, *sinv = ly.sinv + ro * ly.H;
  uint32_t* rec = ctx->rec + v.e0 * ly.recw;
  float* galpha = ly.galpha_dbg ? ly.galpha_dbg + v.e0 * ly.H : nullptr;
  int n_part = 0;
  if (ly.vec && ctx->use_stream && edge_stream_supported(ly.H, ly.D)) {
    if ((phases & 1) && v.nb <= 0) phases &= ~1;
    if (!phases) return GATX_OK;
    if ((phases & 1) && ly.gh_prepared) phases |= 4;
    // pass 2 walks the transposed graph of ALL local edges and reads g_h / rec by local row / edge id: whole range only
    LAUNCHED(launch_edge_backward_stream(gl, ly.H, ly.D, ly.Pl, Pr, a, Hfull, (phases & 2) ? ly.gH : gH, ctx->cdot, score,
                                         mx, sinv, gPr, ctx->gPl, (phases & 2) ? ctx->rec : rec, ctx->part,
                                         ctx->ga_partials, &n_part, galpha, ctx->st, phases));
    if (phases & 1)
      LAUNCHED(launch_reduce_partials(ctx->ga_partials, n_part, ly.F, ctx->grads + ly.a_off, true, ctx->st));
    return GATX_OK;
  }
  if (phases != 3 || !v.whole) return fail(ctx, GATX_ERR_UNSUPPORTED, "layer %d: only the streaming kernels run per block", l);
  if (!ly.vec) {
    LAUNCHED(launch_edge_backward_generic(gl, ly.H, ly.D, ly.Pl, Pr, a, Hfull, gH, score, mx, sinv, gPr, ctx->gPl, ctx->rec,
                                          ctx->ga_partials, &n_part, galpha, ctx->st));
    LAUNCHED(launch_reduce_partials(ctx->ga_partials, n_part, ly.F, ctx->grads + ly.a_off, true, ctx->st));
  } else {
    LAUNCHED(launch_edge_backward_dst(gl, ly.H, ly.D, ly.Pl, Pr, a, Hfull, gH, score, mx, sinv, gPr, ctx->rec,
                                      ctx->ga_partials, &n_part, galpha, ctx->st));
    LAUNCHED(launch_reduce_partials(ctx->ga_partials, n_part, ly.F, ctx->grads + ly.a_off, true, ctx->st));
    LAUNCHED(launch_edge_backward_src(gl, ly.H, ly.D, a, gH, ctx->rec, ctx->gPl, ctx->st));
  }
  return GATX_OK;
}

int do_forward(gatx_ctx* ctx) {
  int rc = ensure_buffers(ctx);
  const unsigned char* mask = ctx->eval_mode ? ctx->eval_mask : ctx->train_mask;
  if (rc) return rc;
  if (!ctx->have_params) return fail(ctx, GATX_ERR_INVALID, "parameters not initialised");
  const float* X = ctx->X0 + (int64_t)ctx->r0 * ctx->ld0;
  const float* X0all = ctx->X0;
  const bool drop = !ctx->eval_mode && ctx->p_drop > 0.f;
  if (drop) ++ctx->drop_step;
  if (!ctx->eval_mode) ctx->fwd_dropped = drop;
  const bool adrop = !ctx->eval_mode && ctx->p_adrop > 0.f;
  if (adrop) ++ctx->adrop_step;
  if (!ctx->eval_mode) ctx->fwd_adropped = adrop;
  const bool p2p_any = ctx->world > 1 && ctx->peers_ready;
  ctx->ev_used = 0;
  if (p2p_any) {
    // every rank has finished ALL its earlier work (the edge passes that read P_l and the gP_l scratch) before the
    // first row of this forward is pushed into its buffers
    stream_after(ctx, ctx->st_comm, ctx->st);
    if ((rc = comm_barrier(ctx))) return rc;
  }
  {
    PhaseTimer t(ctx, PH_GEMM_FWD);
    for (int l = 0; l < ctx->L; ++l) {
      Layer& ly = ctx->layers[l];
      LAUNCHED(launch_pack_weights(ctx->params + ly.w_off, ly.F, ly.I, ly.Wcat, ly.ldk, ly.WcatT, ctx->st));
    }
  }
  const RowView all = row_view(ctx, -1);
  for (int l = 0; l < ctx->L; ++l) {
    Layer& ly = ctx->layers[l];
    const bool replicated = l == 0 && ctx->world > 1;  // layer 0 with the input features on every rank
    // with the peer-memory exchange, P_l / P_r of layers > 0 were produced and pushed block by block while the
    // previous layer's edge pass was still running (below)
    const bool piped_in = l > 0 && halo_p2p(ctx, ly.F);
    if (!piped_in) {
      {
        PhaseTimer t(ctx, PH_GEMM_FWD);
        if (drop) {
          // layer 0 keeps every rank's copy of ALL input rows dropped identically (the mask is a pure function of the
          // global row); deeper layers drop their own rows of the previous layer's output
          const int rows = l == 0 ? ctx->N : ctx->n_rows, row0 = l == 0 ? 0 : ctx->r0;
          if (!ly.Xd) {
            CK(dalloc(&ly.Xd, (size_t)rows * ly.ldx));
            CK(cudaMemsetAsync(ly.Xd, 0, sizeof(float) * (size_t)rows * ly.ldx, ctx->st));  // zero padding columns
          }
          LAUNCHED(launch_dropout(l == 0 ? ctx->X0 : X, ly.ldx, ly.Xd, ly.ldx, rows, ly.I, row0, ctx->p_drop,
                                  ctx->drop_seed, l, ctx->drop_step, ctx->st));
          X = l == 0 ? ly.Xd + (int64_t)ctx->r0 * ly.ldx : ly.Xd;
          X0all = ly.Xd;
        }
        // P_l = X W_l^T, P_r = X W_r^T : the only dense contraction of the forward (EB:303-316)
        if (replicated) {
          rc = gemm_tn_any(ctx, X0all, ly.ldx, ly.Wcat, ly.ldk, ly.Pl, ly.F, ctx->N, ly.F, ly.I);
          if (!rc) rc = gemm_tn_any(ctx, X, ly.ldx, ly.Wcat + (int64_t)ly.F * ly.ldk, ly.ldk, ly.Pr, ly.F, ctx->n_rows, ly.F, ly.I);
        } else {
          rc = gemm_project(ctx, X, ly.ldx, ly, ly.Pl + (int64_t)ctx->r0 * ly.F, ly.Pr, ctx->n_rows);
        }
        if (rc) return rc;
      }
      if (!replicated) {
        rc = comm_allgather_rows(ctx, ly.Pl, ly.F);
        if (rc) return rc;
      }
    }
    const bool pipe_out = l + 1 < ctx->L && halo_p2p(ctx, ctx->layers[l + 1].F);
    if (!pipe_out) {
      PhaseTimer t(ctx, PH_EDGE_FWD);
      if ((rc = fwd_edge(ctx, l, all))) return rc;
      if (ly.Hout != ly.Hfull) LAUNCHED(launch_head_mean(ly.Hfull, ctx->n_rows, ly.H, ly.D, ly.Hout, ctx->st));
    } else {
      // Pipeline over blocks of own destination rows: edge pass of layer l on block b, projection of layer l + 1 on
      // the rows it just produced, then (exchange stream) the push of those projected rows into the peers that gather
      // them -- which overlaps the edge pass of block b + 1.
      Layer& nx = ctx->layers[l + 1];
      const bool blk = blockable(ctx, l);
      if (!blk) {
        PhaseTimer t(ctx, PH_EDGE_FWD);
        if ((rc = fwd_edge(ctx, l, all))) return rc;
      }
      if (drop && !nx.Xd) {
        CK(dalloc(&nx.Xd, (size_t)ctx->n_rows * nx.ldx));
        CK(cudaMemsetAsync(nx.Xd, 0, sizeof(float) * (size_t)ctx->n_rows * nx.ldx, ctx->st));
      }
      const int nblk = (int)ctx->blocks.size();
      for (int b = 0; b < nblk; ++b) {
        const RowView v = row_view(ctx, b);
        if (v.nb <= 0) continue;
        if (blk) {
          PhaseTimer t(ctx, PH_EDGE_FWD);
          if ((rc = fwd_edge(ctx, l, v))) return rc;
        }
        {
          PhaseTimer t(ctx, PH_GEMM_FWD);
          const float* Xb = ly.Hout + (int64_t)v.rb * nx.ldx;
          if (drop) {
            float* Xd = nx.Xd + (int64_t)v.rb * nx.ldx;
            LAUNCHED(launch_dropout(Xb, nx.ldx, Xd, nx.ldx, v.nb, nx.I, ctx->r0 + v.rb, ctx->p_drop, ctx->drop_seed, l + 1,
                                    ctx->drop_step, ctx->st));
            Xb = Xd;
          }
          rc = gemm_project(ctx, Xb, nx.ldx, nx, nx.Pl + (int64_t)(ctx->r0 + v.rb) * nx.F, nx.Pr + (int64_t)v.rb * nx.F, v.nb);
          if (rc) return rc;
        }
        {
          stream_after(ctx, ctx->st_comm, ctx->st);
          CommTimer ct(ctx, 0, (double)ctx->blocks[b].halo_rows * nx.F * 4.0);
          if (ctx->halo_bulk && halo_bulk_supported(nx.F))
            LAUNCHED(launch_halo_push_bulk(nx.Pl + (int64_t)(ctx->r0 + v.rb) * nx.F, ctx->r0 + v.rb, v.nb, nx.F,
                                           ctx->ref_mask + v.rb, ctx->peer_Pl[l + 1], ctx->rank, ctx->st_comm,
                                           halo_cta_slots(ctx->world)));
          else
            LAUNCHED(launch_halo_push(nx.Pl + (int64_t)(ctx->r0 + v.rb) * nx.F, ctx->r0 + v.rb, v.nb, nx.F,
                                      ctx->ref_mask + v.rb, ctx->peer_Pl[l + 1], ctx->rank, ctx->st_comm, halo_cta_slots(ctx->world)));
        }
      }
      if ((rc = comm_barrier(ctx))) return rc;  // every rank's pushes have landed
      compute_waits_comm(ctx);
    }
    X = ly.Hout;
  }
  {
    PhaseTimer t(ctx, PH_HEAD);
    Layer& last = ctx->layers[ctx->L - 1];
    float* gH_out = last.gHout ? last.gHout : last.gH;
    int n_part = 0;
    const float* Wo = ctx->params + ctx->wo_off;
    bool tc_done = false;
    if (ctx->gemm_mode == GATX_GEMM_TF32_TC) {
      // z = H_L W_o^T and dL/dH_L = dz W_o on the tensor cores, softmax / CE / argmax / dz in between
      int n1 = launch_gemm_tc_tn(last.Hout, last.D, Wo, last.D, ctx->z_dbg, ctx->ldc, ctx->n_rows, ctx->C, last.D, false,
                                 ctx->st);
      if (n1 >= 0) {
        ctx->launches += n1;
        LAUNCHED(launch_softmax_ce(ctx->z_dbg, ctx->labels, ctx->n_rows, ctx->C, ctx->ldc, ctx->y, ctx->dz, ctx->pred,
                                   ctx->loss_partials, ctx->correct_partials, &n_part, mask, ctx->st));
        if (!ctx->eval_mode) {
          LAUNCHED(launch_transpose_wo(Wo, ctx->C, last.D, ctx->ldc, ctx->WoT, ctx->st));
          int n2 = launch_gemm_tc_tn(ctx->dz, ctx->ldc, ctx->WoT, ctx->ldc, gH_out, last.D, ctx->n_rows, last.D, ctx->C,
                                     false, ctx->st);
          if (n2 < 0) return fail(ctx, GATX_ERR_UNSUPPORTED, "tensor-core head gradient GEMM rejected its shape");
          ctx->launches += n2;
        }
        tc_done = true;
      }
    }
    if (!tc_done)
      LAUNCHED(launch_head(last.Hout, Wo, ctx->labels, ctx->n_rows, ctx->C, last.D, ctx->ldc, ctx->y, ctx->dz,
                           ctx->keep_debug ? ctx->z_dbg : nullptr, ctx->pred, gH_out, ctx->loss_partials,
                           ctx->correct_partials, &n_part, mask, ctx->st));
    LAUNCHED(launch_loss_finalize(ctx->loss_partials, ctx->correct_partials, n_part, ctx->loss_sum, ctx->correct,
                                  ctx->st));
    if (last.gHout && !ctx->eval_mode)
      LAUNCHED(launch_head_bcast_grad(last.gHout, ctx->n_rows, last.H, last.D, last.gH, ctx->st));
  }
  ctx->fwd_valid = !ctx->eval_mode;
  ctx->last_count = mask ? (ctx->eval_mode ? ctx->eval_count : ctx->train_count) : (int64_t)ctx->N;
  return GATX_OK;
}

// Epilogue arguments that make the input-gradient GEMM of layer l write the pre-activation gradient of layer l - 1 (rows
// [rb, rb + nb) of the own rows) and its segment sums; off (Hout == nullptr) when the layer below does not run the streaming
// kernels (their prep is the only one that can be skipped), with dropout (the mask is applied between GEMM and prep), or
// when GATX_NO_FUSE_PREP is set (A/B).
GemmFuse prep_fuse(const gatx_ctx* ctx, int l, int rb, bool drop) {
  GemmFuse f{};
  static const bool off = getenv("GATX_NO_FUSE_PREP") != nullptr;
  if (off || l <= 0 || drop) return f;
  const Layer& prev = ctx->layers[l - 1];
  if (!(prev.vec && ctx->use_stream && edge_stream_supported(prev.H, prev.D))) return f;
  f.Hout = prev.Hfull + (int64_t)rb * prev.F;
  f.ld = prev.F;
  f.cdot = ctx->cdot;  // rows of the view the following pass 1 is launched on
  f.bias = prev.b_off >= 0 ? ctx->params + prev.b_off : nullptr;
  f.act_slope = ctx->slopes.act;
  f.head_dim = prev.D;
  f.heads = prev.H;
  return f;
}

int do_backward(gatx_ctx* ctx) {
  if (!ctx->have_bufs || !ctx->fwd_valid)
    return fail(ctx, GATX_ERR_INVALID, "gatx_forward must run before gatx_backward (gatx_evaluate does not count)");
  int rc;
  {
    // gW_o += dz^T H_L  (EB:576-581)
    PhaseTimer t(ctx, PH_GEMM_BWD);
    Layer& last = ctx->layers[ctx->L - 1];
    rc = gemm_nt_reduce(ctx, ctx->dz, ctx->ldc, last.Hout, last.D, ctx->grads + ctx->wo_off, last.D, ctx->C, last.D,
                        ctx->n_rows);
    if (rc) return rc;
  }
  const RowView all = row_view(ctx, -1);
  for (Layer& q : ctx->layers) q.gh_prepared = false;
  std::vector<char> p1_done(ctx->L, 0);  // pass 1 of the layer already ran block by block inside the layer above
  for (int l = ctx->L - 1; l >= 0; --l) {
    Layer& ly = ctx->layers[l];
    const bool drop = ctx->fwd_dropped && ly.Xd;
    const float* Xall = drop ? ly.Xd : ctx->X0;  // layer 0: all input rows
    const float* X = l == 0 ? Xall + (int64_t)ctx->r0 * ctx->ld0 : (drop ? ly.Xd : ctx->layers[l - 1].Hout);
    const bool replicated = l == 0 && ctx->world > 1;
    const bool exchanged = ctx->world > 1 && !replicated;
    const bool piped = exchanged && halo_p2p(ctx, ly.F);
    float* gPr = gpr_buf(ctx, l);
    {
      PhaseTimer t(ctx, PH_EDGE_BWD);
      const bool stream = ly.vec && ctx->use_stream && edge_stream_supported(ly.H, ly.D);
      if (stream) {
        if (!p1_done[l] && (rc = bwd_edge(ctx, l, all, 1))) return rc;
        if ((rc = bwd_edge(ctx, l, all, 2))) return rc;
      } else if ((rc = bwd_edge(ctx, l, all, 3))) {
        return rc;
      }
      // gb = column sums of the pre-activation gradient (own rows; the all-reduce of the flat gradients adds the ranks)
      if (ly.b_off >= 0)
        LAUNCHED(launch_colsum(ly.gH, ctx->n_rows, ly.F, ctx->colsum_partials, ctx->grads + ly.b_off, true, ctx->st));
      if (ctx->keep_debug) {
        if (ly.vec) LAUNCHED(launch_unpack_rec(ctx->rec, ctx->E, ly.H, ly.D, ly.alpha_dbg, ly.ge_dbg, ctx->st));
        else LAUNCHED(launch_unpack_rec_generic(ctx->rec, ctx->E, ly.H, ly.alpha_dbg, ly.ge_dbg, ctx->st));
        CK(cudaMemcpyAsync(ly.gPl_dbg, ctx->gPl, sizeof(float) * (size_t)ctx->N * ly.F, cudaMemcpyDeviceToDevice,
                           ctx->st));
        CK(cudaMemcpyAsync(ly.gPr_dbg, gPr, sizeof(float) * (size_t)ctx->n_rows * ly.F, cudaMemcpyDeviceToDevice, ctx->st));
      }
    }
    const float* gPl_own = ctx->gPl + (int64_t)ctx->r0 * ly.F;
    float* gW = ctx->grads + ly.w_off;
    if (!piped) {
      if (exchanged && (rc = comm_reduce_rows(ctx, ctx->gPl, ly.F))) return rc;
      PhaseTimer t(ctx, PH_GEMM_BWD);
      // gW_l = gP_l^T X, gW_r = gP_r^T X  (EB:771-782), W row stride 2I, W_r at column offset I.
      // Replicated layer 0: this rank's PARTIAL gP_l over all sources is contracted with the full X; the
      // all-reduce of the weight gradients completes the sum, so gP_l itself is never exchanged.
      if (replicated)
        rc = gemm_nt_reduce(ctx, ctx->gPl, ly.F, Xall, ly.ldx, gW, 2 * ly.I, ly.F, ly.I, ctx->N);
      else
        rc = gemm_nt_reduce(ctx, gPl_own, ly.F, X, ly.ldx, gW, 2 * ly.I, ly.F, ly.I, ctx->n_rows);
      if (rc) return rc;
      rc = gemm_nt_reduce(ctx, gPr, ly.F, X, ly.ldx, gW + ly.I, 2 * ly.I, ly.F, ly.I, ctx->n_rows);
      if (rc) return rc;
      if (l > 0) {
        // dL/dHout[l-1] = gP_l W_l + gP_r W_r  (EB:859-869); the LReLU derivative of EB:879-893 is
        // applied by the next edge backward when it loads this gradient.
        Layer& prev = ctx->layers[l - 1];
        const GemmFuse fz = prep_fuse(ctx, l, 0, drop);
        rc = gemm_input_grad(ctx, gPl_own, gPr, ly, prev.gH, prev.F, ctx->n_rows, &fz, &prev.gh_prepared);
        if (rc) return rc;
        // gradient w.r.t. the dropped input -> w.r.t. the previous layer's output: same mask, same scale
        if (drop)
          LAUNCHED(launch_dropout(prev.gH, prev.F, prev.gH, prev.F, ctx->n_rows, ly.I, ctx->r0, ctx->p_drop,
                                  ctx->drop_seed, l, ctx->drop_step, ctx->st));
      }
      continue;
    }
    if (!(ctx->halo_bulk && halo_bulk_supported(ly.F))) {
      // Peer-memory exchange, pipelined over blocks of own rows: (exchange stream) the owner pulls and sums the peers'
      // partial gP_l rows of block b; (compute stream) input-gradient GEMM of block b, then prep + pass 1 of the layer
      // BELOW on block b -- while the pull of block b + 1 is in flight.
      stream_after(ctx, ctx->st_comm, ctx->st);
      if ((rc = comm_barrier(ctx))) return rc;  // every rank's partial sums are complete
      {
        PhaseTimer t(ctx, PH_GEMM_BWD);  // needs nothing from the exchange: runs under the barrier and the first pull
        rc = gemm_nt_reduce(ctx, gPr, ly.F, X, ly.ldx, gW + ly.I, 2 * ly.I, ly.F, ly.I, ctx->n_rows);
        if (rc) return rc;
      }
      const bool fuse_p1 = l > 0 && blockable(ctx, l - 1);
      const int nblk = (int)ctx->blocks.size();
      for (int b = 0; b < nblk; ++b) {
        const RowView v = row_view(ctx, b);
        if (v.nb <= 0) continue;
        {
          CommTimer ct(ctx, 1, (double)ctx->blocks[b].halo_rows * ly.F * 4.0);
          LAUNCHED(launch_halo_pull(ctx->gPl + (int64_t)(ctx->r0 + v.rb) * ly.F, ctx->r0 + v.rb, v.nb, ly.F,
                                    ctx->ref_mask + v.rb, ctx->peer_gPl, ctx->rank, ctx->world, ctx->st_comm,
                                    halo_cta_slots(ctx->world)));
        }
        compute_waits_comm(ctx);
        if (l > 0) {
          Layer& prev = ctx->layers[l - 1];
          {
            PhaseTimer t(ctx, PH_GEMM_BWD);
            float* gXb = prev.gH + (int64_t)v.rb * prev.F;
            const GemmFuse fz = prep_fuse(ctx, l, v.rb, drop);  // every block of a layer takes the same path
            rc = gemm_input_grad(ctx, gPl_own + (int64_t)v.rb * ly.F, gPr + (int64_t)v.rb * ly.F, ly, gXb, prev.F, v.nb, &fz,
                                 &prev.gh_prepared);
            if (rc) return rc;
            if (drop)
              LAUNCHED(launch_dropout(gXb, prev.F, gXb, prev.F, v.nb, ly.I, ctx->r0 + v.rb, ctx->p_drop, ctx->drop_seed, l,
                                      ctx->drop_step, ctx->st));
          }
          if (fuse_p1) {
            PhaseTimer t(ctx, PH_EDGE_BWD);
            if ((rc = bwd_edge(ctx, l - 1, v, 1))) return rc;
          }
        }
      }
      if (fuse_p1) p1_done[l - 1] = 1;
      if ((rc = comm_barrier(ctx))) return rc;  // every owner has pulled: the peers may overwrite their gP_l scratch
      {
        PhaseTimer t(ctx, PH_GEMM_BWD);
        rc = gemm_nt_reduce(ctx, gPl_own, ly.F, X, ly.ldx, gW, 2 * ly.I, ly.F, ly.I, ctx->n_rows);
        if (rc) return rc;
      }
      compute_waits_comm(ctx);  // before the next pass 2 writes into the scratch the peers were reading

      continue;
    }
    // Peer-memory exchange, pipelined over blocks of own rows.  Exchange stream: every rank scatters its partial gP_l
    // rows of block b (of every owner) into the owners' staging buffers, then a flag barrier.  Compute stream: the owner
    // adds its own and the staged partials of block b in rank order (local memory), input-gradient GEMM of block b, then
    // prep + pass 1 of the layer BELOW on block b -- while the scatter of block b + 1 is in flight.
    stream_after(ctx, ctx->st_comm, ctx->st);
    if ((rc = comm_barrier(ctx))) return rc;  // every owner is done with the staging contents of the previous exchange
    {
      PhaseTimer t(ctx, PH_GEMM_BWD);  // needs nothing from the exchange: runs under the first scatter
      rc = gemm_nt_reduce(ctx, gPr, ly.F, X, ly.ldx, gW + ly.I, 2 * ly.I, ly.F, ly.I, ctx->n_rows);
      if (rc) return rc;
    }
    const bool fuse_p1 = l > 0 && blockable(ctx, l - 1);
    const int nblk = (int)ctx->blocks.size();
    for (int b = 0; b < nblk; ++b) {
      const RowView v = row_view(ctx, b);
      {
        ScatterPlan plan{};
        for (int p = 0; p < ctx->world; ++p) {
          if (p == ctx->rank) continue;
          const int* ab = ctx->all_blk.data() + (size_t)p * (nblk + 1);
          const int sg = plan.n_seg++;
          plan.row0[sg] = ab[b];
          plan.owner[sg] = p;
          plan.owner_row0[sg] = ctx->bounds[p];
          plan.cum[sg + 1] = plan.cum[sg] + (ab[b + 1] - ab[b]);
          // the owner's staging slot of this rank, viewed with this layer's row pitch
          plan.dst[sg] = ctx->peer_stage.p[p] + (int64_t)ctx->rank * (ctx->bounds[p + 1] - ctx->bounds[p]) * ly.F;
        }
        CommTimer ct(ctx, 1, (double)ctx->scatter_rows[b] * ly.F * 4.0);
        LAUNCHED(launch_halo_scatter_bulk(ctx->gPl, ly.F, ctx->my_ref, plan, ctx->peer_stage, ctx->rank, ctx->st_comm,
                                          halo_cta_slots(ctx->world)));
      }
      if ((rc = comm_barrier(ctx))) return rc;  // block b of every rank's partial rows has landed at its owner
      compute_waits_comm(ctx);
      if (v.nb <= 0) continue;
      {
        PhaseTimer t(ctx, PH_COMM);  // the owner's ordered sum (local memory) belongs to the exchange
        LAUNCHED(launch_halo_sum(ctx->gPl + (int64_t)(ctx->r0 + v.rb) * ly.F, v.rb, v.nb, ctx->n_rows, ly.F,
                                 ctx->ref_mask + v.rb, ctx->stage, ctx->rank, ctx->world, ctx->st));
      }
      if (l > 0) {
        Layer& prev = ctx->layers[l - 1];
        {
          PhaseTimer t(ctx, PH_GEMM_BWD);
          float* gXb = prev.gH + (int64_t)v.rb * prev.F;
          const GemmFuse fz = prep_fuse(ctx, l, v.rb, drop);  // every block of a layer takes the same path
          rc = gemm_input_grad(ctx, gPl_own + (int64_t)v.rb * ly.F, gPr + (int64_t)v.rb * ly.F, ly, gXb, prev.F, v.nb, &fz,
                               &prev.gh_prepared);
          if (rc) return rc;
          if (drop)
            LAUNCHED(launch_dropout(gXb, prev.F, gXb, prev.F, v.nb, ly.I, ctx->r0 + v.rb, ctx->p_drop, ctx->drop_seed, l,
                                    ctx->drop_step, ctx->st));
        }
        if (fuse_p1) {
          PhaseTimer t(ctx, PH_EDGE_BWD);
          if ((rc = bwd_edge(ctx, l - 1, v, 1))) return rc;
        }
      }
    }
    if (fuse_p1) p1_done[l - 1] = 1;
    {
      PhaseTimer t(ctx, PH_GEMM_BWD);
      rc = gemm_nt_reduce(ctx, gPl_own, ly.F, X, ly.ldx, gW, 2 * ly.I, ly.F, ly.I, ctx->n_rows);
      if (rc) return rc;
    }
    }
  return GATX_OK;
}

int do_step(gatx_ctx* ctx, int t) {
  if (!ctx->have_bufs) return fail(ctx, GATX_ERR_INVALID, "nothing to update");
  if (ctx->world > 1) {
    if (!ctx->comm) return fail(ctx, GATX_ERR_INVALID, "world > 1 but gatx_comm_init was not called");
    PhaseTimer tm(ctx, PH_COMM);
    NK(g_nccl.AllReduce(ctx->grads, ctx->grads, (size_t)ctx->n_params, ncclFloat, ncclSum, ctx->comm, ctx->st));
  }
  PhaseTimer tm(ctx, PH_OPT);
  LAUNCHED(launch_optimizer(ctx->params, ctx->grads, ctx->adam_m, ctx->adam_v, ctx->bias_begin, ctx->grp, ctx->clip != 0,
                            ctx->optimizer, ctx->lr, ctx->b1, ctx->b2, t, ctx->norm_partials, ctx->st));
  if (ctx->bias_end > ctx->bias_begin) {  // the biases are a clip group of their own
    const int64_t o = ctx->bias_begin, n = ctx->bias_end - ctx->bias_begin;
    OptimGroups one{};
    one.begin[0] = 0;
    one.end[0] = one.begin[1] = one.end[1] = one.begin[2] = one.end[2] = n;
    LAUNCHED(launch_optimizer(ctx->params + o, ctx->grads + o, ctx->adam_m + o, ctx->adam_v + o, n, one, ctx->clip != 0,
                              ctx->optimizer, ctx->lr, ctx->b1, ctx->b2, t, ctx->norm_partials + 3 * kOptimBlocks,
                              ctx->st));
  }
  return GATX_OK;
}

__global__ void pack_loss_kernel(const double* loss_sum, const long long* correct, double* out2) {
  out2[0] = *loss_sum;
  out2[1] = (double)*correct;
}

int read_loss(gatx_ctx* ctx, float* avg_loss, float* accuracy, int64_t count) {
  pack_loss_kernel<<<1, 1, 0, ctx->st>>>(ctx->loss_sum, ctx->correct, ctx->red2);
  ctx->launches += 1;
  if (ctx->world > 1) {
    if (!ctx->comm) return fail(ctx, GATX_ERR_INVALID, "world > 1 but gatx_comm_init was not called");
    NK(g_nccl.AllReduce(ctx->red2, ctx->red2, 2, ncclDouble, ncclSum, ctx->comm, ctx->st));
  }
  double h[2];
  CK(cudaMemcpyAsync(h, ctx->red2, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  const double den = (double)(count > 0 ? count : 1);
  if (avg_loss) *avg_loss = (float)(h[0] / den);  // EB:544 (count = N without a mask)
  if (accuracy) *accuracy = (float)(h[1] / den);  // EB:546
  return GATX_OK;
}

// ---- CUDA-graph replay of the epoch ----------------------------------------------------------------------------
// Small graphs (cora / pubmed class) run ~50 kernels of a few microseconds each per epoch: the epoch is bound by
// launch latency, not by the GPU.  Forward + backward are captured once per (buffers, graph, mask) generation and
// replayed with one cudaGraphLaunch.  Single rank only (the multi-rank epoch interleaves NCCL and peer-memory
// barriers), not while per-phase timing is on (event records change the launch sequence).
constexpr int64_t kGraphAutoMaxEdges = 8 << 20;  // beyond this the kernels are long enough to hide the launches

bool epoch_graph_wanted(const gatx_ctx* ctx) {
  // dropout: the step counter is a kernel argument that changes every epoch
  if (ctx->world != 1 || ctx->timing || ctx->graph_mode == 0 || ctx->p_drop > 0.f || ctx->p_adrop > 0.f) return false;
  if (ctx->graph_mode == 1) return true;
  static const int env = [] {
    const char* e = getenv("GATX_CUDA_GRAPH");
    return e ? atoi(e) : -1;
  }();
  if (env == 0) return false;
  return env == 1 || ctx->E <= kGraphAutoMaxEdges;
}

// Captures do_forward + do_backward into ctx->epoch_exec.  A capture that fails leaves epoch_exec null and turns the
// replay off for this context (the eager path then runs the epoch); a genuine launch error is reported as usual.
int capture_epoch(gatx_ctx* ctx) {
  if (ctx->epoch_exec) {
    cudaGraphExecDestroy(ctx->epoch_exec);
    ctx->epoch_exec = nullptr;
  }
  CK(cudaStreamSynchronize(ctx->st));
  const int64_t before = ctx->launches;
  if (cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    ctx->graph_mode = 0;
    return GATX_OK;
  }
  int rc = do_forward(ctx);
  if (!rc) rc = do_backward(ctx);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->st, &graph);
  const int64_t n = ctx->launches - before;
  ctx->launches = before;
  ctx->fwd_valid = false;
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess || !graph || cudaGraphInstantiate(&ctx->epoch_exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    ctx->epoch_exec = nullptr;
    ctx->graph_mode = 0;
    return GATX_OK;
  }
  cudaGraphDestroy(graph);
  ctx->epoch_exec_gen = ctx->gen;
  ctx->epoch_exec_launches = n;
  ctx->epoch_exec_count = ctx->last_count;
  return GATX_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

const char* gatx_version(void) { return "gatx 0.2 (sm_100a)"; }

int gatx_device_count(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int usable = 0;
  for (int d = 0; d < ndev; ++d) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) != cudaSuccess || prop.major < 10) break;  // contexts use ordinals 0..n-1
    ++usable;
  }
  return usable;
}

int gatx_create(gatx_ctx** out, const gatx_config* cfg) {
  if (!out || !cfg || cfg->num_layers <= 0 || !cfg->heads || !cfg->outdims) return GATX_ERR_INVALID;
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return GATX_ERR_INVALID;
  if (cfg->optimizer == GATX_OPT_ADAM &&
      !(cfg->beta1 > 0.f && cfg->beta1 < 1.f && cfg->beta2 > 0.f && cfg->beta2 < 1.f))
    return GATX_ERR_INVALID;  // EB:1011-1015
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) return GATX_ERR_CUDA;
  if (cudaSetDevice(cfg->device) != cudaSuccess) return GATX_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major < 10) return GATX_ERR_CUDA;
  gatx_ctx* c = new gatx_ctx();
  c->L = cfg->num_layers;
  c->heads.assign(cfg->heads, cfg->heads + c->L);
  c->outdims.assign(cfg->outdims, cfg->outdims + c->L);
  c->optimizer = cfg->optimizer; c->clip = cfg->clip; c->device = cfg->device;
  c->gemm_mode = cfg->gemm_mode; c->keep_debug = cfg->keep_debug; c->rank = cfg->rank; c->world = cfg->world;
  c->lr = cfg->lr; c->b1 = cfg->beta1; c->b2 = cfg->beta2;
  c->layers.resize(c->L);
  for (int l = 0; l < c->L; ++l)
    if (c->heads[l] <= 0 || c->outdims[l] <= 0) {
      delete c;
      return GATX_ERR_INVALID;
    }
  if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return GATX_ERR_CUDA;
  }
  if (c->world > 1) {
    // the exchange stream: highest priority, so that its (small) kernels get the first CTA slots the edge pass frees
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&c->st_comm, cudaStreamNonBlocking, hi) != cudaSuccess) {
      cudaStreamDestroy(c->st);
      delete c;
      return GATX_ERR_CUDA;
    }
    const char* hm = getenv("GATX_HALO_MODE");
    c->halo_bulk = hm && std::string(hm) == "bulk";
  }
  *out = c;
  return GATX_OK;
}

void gatx_destroy(gatx_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  if (ctx->st_comm) cudaStreamSynchronize(ctx->st_comm);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  if (ctx->epoch_exec) cudaGraphExecDestroy(ctx->epoch_exec);
  free_bufs(ctx);
  free_graph(ctx);
  dfree(ctx->X0);
  dfree(ctx->labels);
  dfree(ctx->train_mask);
  dfree(ctx->eval_mask);
  for (auto& s : ctx->spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  if (ctx->sw_a) {
    cudaEventDestroy(ctx->sw_a);
    cudaEventDestroy(ctx->sw_b);
  }
  for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto& s : ctx->comm_spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  if (ctx->st_comm) cudaStreamDestroy(ctx->st_comm);
  cudaStreamDestroy(ctx->st);
  delete ctx;
}

const char* gatx_last_error(const gatx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gatx_partition_rows(int32_t num_nodes, const int32_t* row_ptr, int32_t world, int32_t* bounds) {
  if (num_nodes < 0 || !row_ptr || world < 1 || !bounds) return GATX_ERR_INVALID;
  // rank r owns rows [bounds[r], bounds[r+1]); bounds[r] = first row with row_ptr >= floor(r*E/world)
  const int64_t E = row_ptr[num_nodes];
  bounds[0] = 0;
  int i = 0;
  for (int r = 1; r < world; ++r) {
    const int64_t target = E * (int64_t)r / (int64_t)world;
    while (i < num_nodes && (int64_t)row_ptr[i] < target) ++i;
    bounds[r] = i;
  }
  bounds[world] = num_nodes;
  return GATX_OK;
}

int gatx_row_blocks(int32_t num_nodes, const int32_t* row_ptr, int32_t world, int32_t num_blocks, int32_t* out) {
  if (num_nodes < 0 || !row_ptr || world < 1 || num_blocks < 1 || !out) return GATX_ERR_INVALID;
  std::vector<int32_t> bounds((size_t)world + 1);
  gatx_partition_rows(num_nodes, row_ptr, world, bounds.data());
  const int K = num_blocks;
  for (int p = 0; p < world; ++p) {
    const int b0 = bounds[p], b1 = bounds[p + 1];
    const int64_t base = row_ptr[b0], ep = (int64_t)row_ptr[b1] - base;
    int32_t* ab = out + (size_t)p * (K + 1);
    ab[0] = b0;
    ab[K] = b1;
    for (int k = 1, i = b0; k < K; ++k) {  // first own row whose edge offset reaches k E_p / K
      const int64_t target = ep * (int64_t)k / K;
      while (i < b1 && (int64_t)row_ptr[i] - base < target) ++i;
      ab[k] = i;
    }
  }
  return GATX_OK;
}

int gatx_set_graph_csr(gatx_ctx* ctx, int32_t N, int64_t E, const int32_t* row_ptr, const int32_t* col_idx) {
  if (!ctx || N <= 0 || E < 0 || !row_ptr || (E > 0 && !col_idx)) return fail(ctx, GATX_ERR_INVALID, "bad graph");
  if (row_ptr[0] != 0 || (int64_t)row_ptr[N] != E) return fail(ctx, GATX_ERR_INVALID, "row_ptr[N] != num_edges");
  CK(cudaSetDevice(ctx->device));
  free_bufs(ctx);  // every buffer is sized by the graph
  free_graph(ctx);
  dfree(ctx->X0);
  dfree(ctx->labels);
  dfree(ctx->train_mask);  // masks are per node: a new graph drops them
  dfree(ctx->eval_mask);
  ctx->fwd_valid = false;
  ctx->have_feat = ctx->have_labels = false;
  ctx->N = N;
  ctx->Eg = E;
  ctx->bounds.resize(ctx->world + 1);
  gatx_partition_rows(N, row_ptr, ctx->world, ctx->bounds.data());
  ctx->r0 = ctx->bounds[ctx->rank];
  ctx->r1 = ctx->bounds[ctx->rank + 1];
  ctx->n_rows = ctx->r1 - ctx->r0;
  const int64_t e0 = row_ptr[ctx->r0], e1 = row_ptr[ctx->r1];
  ctx->E = e1 - e0;
  ctx->edge0 = e0;
  int maxdeg = 0;
  std::vector<int> local_ptr(ctx->n_rows + 1), heavy;
  std::vector<uint16_t> ref_own;  // ref_mask of the own rows (host copy, for the per-block halo counts)
  std::vector<unsigned char> my_ref_host;  // [N] this rank's edges gather the source
  for (int i = 0; i < N; ++i) {
    if (row_ptr[i + 1] < row_ptr[i]) return fail(ctx, GATX_ERR_INVALID, "row_ptr not monotone at %d", i);
    const int d = row_ptr[i + 1] - row_ptr[i];
    if (d > maxdeg) maxdeg = d;  // EB:89-99
  }
  ctx->max_degree = maxdeg;
  if (ctx->world > 1 && ctx->world <= kMaxPeers) {
    // which ranks' edge slices reference each source (integer work, identical on every rank): the halo lists
    std::vector<uint16_t> ref((size_t)N, 0);
    for (int p = 0; p < ctx->world; ++p) {
      const uint16_t bit = (uint16_t)(1u << p);
      const int64_t a = row_ptr[ctx->bounds[p]], b = row_ptr[ctx->bounds[p + 1]];
      for (int64_t e = a; e < b; ++e) {
        const int sidx = col_idx[e];
        if (sidx >= 0 && sidx < N) ref[sidx] |= bit;
      }
    }
    int64_t halo = 0;
    for (int i = ctx->r0; i < ctx->r1; ++i) halo += __builtin_popcount((unsigned)(ref[i] & ~(1u << ctx->rank)));
    ctx->halo_rows = halo;
    ref_own.assign(ref.begin() + ctx->r0, ref.begin() + ctx->r1);
    my_ref_host.resize((size_t)N);
    for (int i = 0; i < N; ++i) my_ref_host[i] = (unsigned char)((ref[i] >> ctx->rank) & 1u);
    CK(dalloc(&ctx->ref_mask, (size_t)ctx->n_rows));
    if (ctx->n_rows)
      CK(cudaMemcpyAsync(ctx->ref_mask, ref.data() + ctx->r0, sizeof(uint16_t) * (size_t)ctx->n_rows,
                         cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));  // `ref` goes out of scope
  }
  for (int i = 0; i <= ctx->n_rows; ++i) local_ptr[i] = (int)(row_ptr[ctx->r0 + i] - e0);
  for (int i = 0; i < ctx->n_rows; ++i)
    if (local_ptr[i + 1] - local_ptr[i] > kHeavyDeg) heavy.push_back(i);
  for (int64_t e = e0; e < e1; ++e)
    if (col_idx[e] < 0 || col_idx[e] >= N) return fail(ctx, GATX_ERR_INVALID, "col_idx[%lld] out of range", (long long)e);
  CK(dalloc(&ctx->row_ptr, (size_t)ctx->n_rows + 1));
  CK(dalloc(&ctx->col_idx, (size_t)ctx->E));
  CK(dalloc(&ctx->coo_src, (size_t)ctx->E));
  CK(dalloc(&ctx->coo_dst, (size_t)ctx->E));
  CK(dalloc(&ctx->in_deg, (size_t)ctx->n_rows));
  CK(dalloc(&ctx->csc_ptr, (size_t)N + 1));
  CK(dalloc(&ctx->csc_dst, (size_t)ctx->E));
  CK(dalloc(&ctx->csc_eid, (size_t)ctx->E));
  CK(cudaMemcpyAsync(ctx->row_ptr, local_ptr.data(), sizeof(int) * local_ptr.size(), cudaMemcpyHostToDevice, ctx->st));
  if (ctx->E)
    CK(cudaMemcpyAsync(ctx->col_idx, col_idx + e0, sizeof(int) * (size_t)ctx->E, cudaMemcpyHostToDevice, ctx->st));
  ctx->n_heavy_rows = (int)heavy.size();
  CK(dalloc(&ctx->heavy_rows, heavy.size()));
  if (!heavy.empty())
    CK(cudaMemcpyAsync(ctx->heavy_rows, heavy.data(), sizeof(int) * heavy.size(), cudaMemcpyHostToDevice, ctx->st));
  LAUNCHED(launch_csr_to_coo(ctx->row_ptr, ctx->col_idx, ctx->coo_src, ctx->coo_dst, ctx->in_deg, ctx->n_rows, ctx->st));
  int n = build_csc(ctx->col_idx, ctx->coo_dst, ctx->E, N, ctx->csc_ptr, ctx->csc_dst, ctx->csc_eid, ctx->st);
  if (n < 0) return fail(ctx, GATX_ERR_CUDA, "build_csc: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches += n;
  // edge-balanced chunking for the streaming kernels
  if (const char* ev = getenv("GATX_CHUNK")) {
    const int t = atoi(ev);
    if (t >= 32 && t <= (1 << 20)) ctx->chunk_T = t;
  }
  ctx->use_stream = getenv("GATX_NO_STREAM") == nullptr;
  ctx->n_chunks = (int)((ctx->E + ctx->chunk_T - 1) / ctx->chunk_T);
  CK(dalloc(&ctx->chunk_row, (size_t)ctx->n_chunks + 1));
  CK(dalloc(&ctx->chunk_src, (size_t)ctx->n_chunks + 1));
  LAUNCHED(launch_chunk_rows(ctx->row_ptr, ctx->n_rows, ctx->E, ctx->chunk_T, ctx->n_chunks, ctx->chunk_row, ctx->st));
  LAUNCHED(launch_chunk_rows(ctx->csc_ptr, N, ctx->E, ctx->chunk_T, ctx->n_chunks, ctx->chunk_src, ctx->st));
  if (ctx->world > 1) {
    // Blocks of own rows for the pipelined exchange (edge-balanced like the rank partition itself).  A block must still
    // fill the GPU: at least 8 chunks per SM on EVERY rank (the block count is part of the exchange protocol, so it is
    // derived from the global row_ptr), unless GATX_HALO_BLOCKS forces a count (tests).
    int K = 4;
    bool forced = false;
    if (const char* ev = getenv("GATX_HALO_BLOCKS")) {
      const int k = atoi(ev);
      if (k >= 1 && k <= 16) { K = k; forced = true; }
    }
    int64_t min_chunks = INT64_MAX;
    int min_rows = N;
    for (int p = 0; p < ctx->world; ++p) {
      const int64_t ep = (int64_t)row_ptr[ctx->bounds[p + 1]] - row_ptr[ctx->bounds[p]];
      min_chunks = std::min<int64_t>(min_chunks, (ep + ctx->chunk_T - 1) / ctx->chunk_T);
      min_rows = std::min(min_rows, ctx->bounds[p + 1] - ctx->bounds[p]);
    }
    while (!forced && K > 1 && min_chunks / K < kNumSMs * 8) --K;
    if (K > min_rows) K = min_rows > 0 ? min_rows : 1;
    // every rank's blocks (global rows): rank p's block k starts at the first own row whose edge offset >= k E_p / K
    ctx->all_blk.assign((size_t)ctx->world * (K + 1), 0);
    gatx_row_blocks(N, row_ptr, ctx->world, K, ctx->all_blk.data());
    ctx->blocks.assign(K, gatx_ctx::RowBlock{});
    const int* rbg = ctx->all_blk.data() + (size_t)ctx->rank * (K + 1);
    std::vector<int> brp;
    int64_t total_chunks = 0;
    for (int k = 0; k < K; ++k) {
      gatx_ctx::RowBlock& B = ctx->blocks[k];
      const int lo = rbg[k] - ctx->r0, hi = rbg[k + 1] - ctx->r0;
      B.r0 = lo;
      B.n_rows = hi - lo;
      B.e0 = local_ptr[lo];
      B.E = local_ptr[hi] - B.e0;
      B.n_chunks = (int)((B.E + ctx->chunk_T - 1) / ctx->chunk_T);
      total_chunks += B.n_chunks + 1;
      for (int i = lo; i <= hi; ++i) brp.push_back((int)(local_ptr[i] - B.e0));
      for (int i = lo; i < hi && !ref_own.empty(); ++i)
        B.halo_rows += __builtin_popcount((unsigned)(ref_own[i] & ~(1u << ctx->rank)));
    }
    if (!my_ref_host.empty()) {
      // rows this rank sends in the backward exchange of block k: referenced sources inside block k of every other owner
      ctx->scatter_rows.assign(K, 0);
      for (int p = 0; p < ctx->world; ++p) {
        if (p == ctx->rank) continue;
        const int* ab = ctx->all_blk.data() + (size_t)p * (K + 1);
        for (int k = 0; k < K; ++k)
          for (int i = ab[k]; i < ab[k + 1]; ++i) ctx->scatter_rows[k] += my_ref_host[i];
      }
      CK(dalloc(&ctx->my_ref, (size_t)N));
      CK(cudaMemcpyAsync(ctx->my_ref, my_ref_host.data(), (size_t)N, cudaMemcpyHostToDevice, ctx->st));
      CK(cudaStreamSynchronize(ctx->st));
    }
    if (K > 1) {
      CK(dalloc(&ctx->blk_row_ptr, brp.size()));
      CK(dalloc(&ctx->blk_chunk_row, (size_t)total_chunks));
      CK(cudaMemcpyAsync(ctx->blk_row_ptr, brp.data(), sizeof(int) * brp.size(), cudaMemcpyHostToDevice, ctx->st));
      size_t po = 0, co = 0;
      for (int k = 0; k < K; ++k) {
        gatx_ctx::RowBlock& B = ctx->blocks[k];
        B.row_ptr = ctx->blk_row_ptr + po;
        B.chunk_row = ctx->blk_chunk_row + co;
        po += (size_t)B.n_rows + 1;
        co += (size_t)B.n_chunks + 1;
        LAUNCHED(launch_chunk_rows(B.row_ptr, B.n_rows, B.E, ctx->chunk_T, B.n_chunks, B.chunk_row, ctx->st));
      }
      CK(cudaStreamSynchronize(ctx->st));  // `brp` goes out of scope
    } else {
      ctx->blocks[0].row_ptr = ctx->row_ptr;
      ctx->blocks[0].chunk_row = ctx->chunk_row;
    }
  }
  // sources with a heavy out-degree (CTA-per-row in the source-major backward pass)
  std::vector<int> cptr((size_t)N + 1), heavy_s;
  CK(cudaMemcpyAsync(cptr.data(), ctx->csc_ptr, sizeof(int) * cptr.size(), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  for (int j = 0; j < N; ++j)
    if (cptr[j + 1] - cptr[j] > kHeavyDeg) heavy_s.push_back(j);
  ctx->n_heavy_srcs = (int)heavy_s.size();
  CK(dalloc(&ctx->heavy_srcs, heavy_s.size()));
  if (!heavy_s.empty())
    CK(cudaMemcpyAsync(ctx->heavy_srcs, heavy_s.data(), sizeof(int) * heavy_s.size(), cudaMemcpyHostToDevice, ctx->st));
  {
    // L2 residency hints for the gathers (edge_stream.cu): on a power-law graph a few percent of the nodes take a
    // large share of the gathers (products shape: the 50 000 most referenced sources, 2 % of the nodes, are 25 % of
    // all gathers).  Mark as many of the most-gathered nodes as fit a budget of the 126 MB L2; their rows are
    // fetched evict_last, all other traffic evict_first.  GATX_HOT_MB sets the budget (0 = off).
    int wideF = 0, narrowF = 0;
    for (int l = 0; l < ctx->L; ++l) {
      const int F = ctx->heads[l] * ctx->outdims[l];
      if (!edge_stream_supported(ctx->heads[l], ctx->outdims[l]) || F < 256) continue;  // 128-float rows: no hints
      if (F > wideF) wideF = F;
      if (narrowF == 0 || F < narrowF) narrowF = F;
    }
    double budget_mb = 96.0;
    if (const char* ev = getenv("GATX_HOT_MB")) budget_mb = atof(ev);
    if (wideF > 0 && budget_mb > 0.0 && N < (1 << 30) && ctx->E > 0) {
      auto threshold = [](std::vector<int> deg, int64_t k) {  // k-th largest degree, at least 2 (a row used once is not hot)
        if (k <= 0) return 0x7fffffff;
        if (k >= (int64_t)deg.size()) return 2;
        std::nth_element(deg.begin(), deg.begin() + (k - 1), deg.end(), [](int a, int b) { return a > b; });
        return deg[k - 1] > 2 ? deg[k - 1] : 2;
      };
      std::vector<int> outdeg((size_t)N), indeg((size_t)ctx->n_rows);
      for (int j = 0; j < N; ++j) outdeg[j] = cptr[j + 1] - cptr[j];
      for (int i = 0; i < ctx->n_rows; ++i) indeg[i] = local_ptr[i + 1] - local_ptr[i];
      const int64_t k_wide = (int64_t)(budget_mb * 1e6 / (4.0 * wideF)), k_narrow = (int64_t)(budget_mb * 1e6 / (4.0 * narrowF));
      const int to_w = threshold(outdeg, k_wide), to_n = threshold(outdeg, k_narrow);
      const int ti_w = threshold(indeg, k_wide), ti_n = threshold(indeg, k_narrow);
      CK(dalloc(&ctx->col_idx_hot, (size_t)ctx->E));
      CK(dalloc(&ctx->csc_dst_hot, (size_t)ctx->E));
      LAUNCHED(launch_mark_hot(ctx->col_idx, ctx->csc_ptr, ctx->E, to_w, to_n, ctx->col_idx_hot, ctx->st));
      LAUNCHED(launch_mark_hot(ctx->csc_dst, ctx->row_ptr, ctx->E, ti_w, ti_n, ctx->csc_dst_hot, ctx->st));
      ctx->hot_wide_F = wideF;
    }
  }
  CK(cudaStreamSynchronize(ctx->st));
  ctx->have_graph = true;
  ++ctx->gen;
  return GATX_OK;
}

int gatx_set_features(gatx_ctx* ctx, const float* X, int32_t in_dim) {
  if (!ctx || !X || in_dim <= 0) return fail(ctx, GATX_ERR_INVALID, "bad features");
  if (!ctx->have_graph) return fail(ctx, GATX_ERR_INVALID, "set the graph before the features");
  CK(cudaSetDevice(ctx->device));
  const int ld = (in_dim + 3) / 4 * 4;  // 16-byte row pitch for TMA / 128-bit loads, zero padded
  if (!ctx->X0 || ctx->I0 != in_dim) {
    if (ctx->have_bufs) free_bufs(ctx);
    // every rank keeps ALL rows of the (constant) input features: layer 0 can then project every source locally
    // and contract its partial gP_l with the full X, so it needs no all-gather and no reduce (see do_forward)
    CK(dalloc(&ctx->X0, (size_t)ctx->N * ld));
    CK(cudaMemsetAsync(ctx->X0, 0, sizeof(float) * (size_t)ctx->N * ld, ctx->st));
    ctx->I0 = in_dim;
    ctx->ld0 = ld;
    ++ctx->gen;
  }
  // With a communicator every rank copies only ITS rows over PCIe and the row blocks are all-gathered over NVLink
  // (the call is then collective: every rank must make it); without one each rank uploads the whole matrix.
  const bool scatter = ctx->world > 1 && ctx->comm != nullptr;
  const int64_t first = scatter ? ctx->r0 : 0, count = scatter ? ctx->n_rows : ctx->N;
  if (count > 0) {
    float* dst = ctx->X0 + first * ld;
    const float* src = X + first * in_dim;
    if (ld == in_dim)  // no padding: one contiguous DMA
      CK(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)count * in_dim, cudaMemcpyHostToDevice, ctx->st));
    else
      CK(cudaMemcpy2DAsync(dst, sizeof(float) * ld, src, sizeof(float) * in_dim, sizeof(float) * in_dim, count,
                           cudaMemcpyHostToDevice, ctx->st));
  }
  if (scatter) {
    int rc = nccl_allgather_rows(ctx, ctx->X0, ld);
    if (rc) return rc;
  }
  ctx->have_feat = true;
  return GATX_OK;
}

int gatx_set_labels(gatx_ctx* ctx, const int32_t* labels, int32_t num_classes) {
  if (!ctx || !labels) return fail(ctx, GATX_ERR_INVALID, "bad labels");
  if (!ctx->have_graph) return fail(ctx, GATX_ERR_INVALID, "set the graph before the labels");
  CK(cudaSetDevice(ctx->device));
  int C = num_classes;
  if (C <= 0) {
    int mx = labels[0];
    for (int i = 1; i < ctx->N; ++i) mx = labels[i] > mx ? labels[i] : mx;  // EB:1106-1107
    C = mx + 1;
  }
  // the loss / gradient kernels index a row of C class scores with the label (EB:524, EB:572): refuse anything outside
  for (int i = 0; i < ctx->N; ++i)
    if (labels[i] < 0 || labels[i] >= C)
      return fail(ctx, GATX_ERR_INVALID, "label %d of node %d outside [0, %d)", labels[i], i, C);
  if (ctx->have_bufs && C != ctx->C) free_bufs(ctx);
  ctx->C = C;
  if (!ctx->labels) {
    CK(dalloc(&ctx->labels, (size_t)ctx->n_rows));
    ++ctx->gen;
  }
  if (ctx->n_rows)
    CK(cudaMemcpyAsync(ctx->labels, labels + ctx->r0, sizeof(int) * (size_t)ctx->n_rows, cudaMemcpyHostToDevice,
                       ctx->st));
  ctx->have_labels = true;
  return GATX_OK;
}

// Uploads this rank's rows of a global 0/1 node mask (or drops the mask when `mask` is NULL).
static int upload_mask(gatx_ctx* ctx, const uint8_t* mask, unsigned char** dev, int64_t* count) {
  if (!ctx->have_graph) return fail(ctx, GATX_ERR_INVALID, "set the graph before a mask");
  CK(cudaSetDevice(ctx->device));
  if (!mask) {
    dfree(*dev);
    *count = ctx->N;
    return GATX_OK;
  }
  int64_t cnt = 0;
  for (int i = 0; i < ctx->N; ++i) cnt += mask[i] != 0;
  if (!*dev) CK(dalloc(dev, (size_t)ctx->n_rows));
  if (ctx->n_rows) {
    CK(cudaMemcpyAsync(*dev, mask + ctx->r0, (size_t)ctx->n_rows, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
  }
  *count = cnt;
  return GATX_OK;
}

int gatx_set_train_mask(gatx_ctx* ctx, const uint8_t* mask) {
  if (!ctx) return GATX_ERR_INVALID;
  ++ctx->gen;  // the mask pointer and the loss denominator are baked into a captured epoch
  return upload_mask(ctx, mask, &ctx->train_mask, &ctx->train_count);
}

int gatx_evaluate(gatx_ctx* ctx, const uint8_t* mask, float* avg_loss, float* accuracy) {
  if (!ctx) return GATX_ERR_INVALID;
  int rc = upload_mask(ctx, mask, &ctx->eval_mask, &ctx->eval_count);
  if (rc) return rc;
  ctx->spans_used = 0;
  ctx->eval_mode = true;
  rc = do_forward(ctx);
  ctx->eval_mode = false;
  if (rc) return rc;
  CK(cudaGetLastError());
  return read_loss(ctx, avg_loss, accuracy, ctx->last_count);
}

int gatx_graph_info(gatx_ctx* ctx, int32_t* max_degree, int32_t* num_classes, int32_t* row_begin, int32_t* row_end) {
  if (!ctx || !ctx->have_graph) return fail(ctx, GATX_ERR_INVALID, "no graph");
  if (max_degree) *max_degree = ctx->max_degree;
  if (num_classes) *num_classes = ctx->C;
  if (row_begin) *row_begin = ctx->r0;
  if (row_end) *row_end = ctx->r1;
  return GATX_OK;
}

int gatx_init_params(gatx_ctx* ctx, uint64_t seed) {
  if (!ctx) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  for (int l = 0; l < ctx->L; ++l) {
    Layer& ly = ctx->layers[l];
    const float limit = sqrtf(6.0f / (float)(2 * ly.I + ly.D));  // EB:208
    LAUNCHED(launch_philox_uniform(ctx->params + ly.w_off, (int64_t)ly.F * 2 * ly.I, limit, seed, 2 * l, ctx->st));
    LAUNCHED(launch_philox_uniform(ctx->params + ly.a_off, ly.F, limit, seed, 2 * l + 1, ctx->st));
  }
  const int DL = ctx->outdims[ctx->L - 1];
  const float limit = sqrtf(6.0f / (float)(ctx->C + DL));  // EB:236
  LAUNCHED(launch_philox_uniform(ctx->params + ctx->wo_off, (int64_t)ctx->C * DL, limit, seed, 1000, ctx->st));
  if (ctx->bias_end > ctx->bias_begin)  // biases start at zero
    CK(cudaMemsetAsync(ctx->params + ctx->bias_begin, 0, sizeof(float) * (size_t)(ctx->bias_end - ctx->bias_begin), ctx->st));
  ctx->have_params = true;
  return GATX_OK;
}

int gatx_set_params(gatx_ctx* ctx, int32_t layer, const float* W, const float* a) {
  if (!ctx || layer < 0 || layer >= ctx->L || !W || !a) return fail(ctx, GATX_ERR_INVALID, "bad set_params");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  Layer& ly = ctx->layers[layer];
  CK(cudaMemcpyAsync(ctx->params + ly.w_off, W, sizeof(float) * (size_t)ly.F * 2 * ly.I, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(ctx->params + ly.a_off, a, sizeof(float) * (size_t)ly.F, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->have_params = true;
  return GATX_OK;
}

int gatx_set_wo(gatx_ctx* ctx, const float* Wo) {
  if (!ctx || !Wo) return fail(ctx, GATX_ERR_INVALID, "bad set_wo");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->params + ctx->wo_off, Wo, sizeof(float) * (size_t)ctx->C * ctx->outdims[ctx->L - 1],
                     cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->have_params = true;
  return GATX_OK;
}

int gatx_forward(gatx_ctx* ctx) {
  if (!ctx) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  ctx->spans_used = 0;
  int rc = do_forward(ctx);
  if (rc) return rc;
  CK(cudaGetLastError());
  return GATX_OK;
}

int gatx_loss_acc(gatx_ctx* ctx, float* avg_loss, float* accuracy) {
  if (!ctx || !ctx->have_bufs) return fail(ctx, GATX_ERR_INVALID, "forward must run first");
  CK(cudaSetDevice(ctx->device));
  return read_loss(ctx, avg_loss, accuracy, ctx->last_count);
}

int gatx_backward(gatx_ctx* ctx) {
  if (!ctx) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = do_backward(ctx);
  if (rc) return rc;
  CK(cudaGetLastError());
  return GATX_OK;
}

int gatx_step(gatx_ctx* ctx, int32_t t) {
  if (!ctx || t < 1) return fail(ctx, GATX_ERR_INVALID, "step index is the 1-based epoch");
  CK(cudaSetDevice(ctx->device));
  int rc = do_step(ctx, t);
  if (rc) return rc;
  CK(cudaGetLastError());
  return GATX_OK;
}

int gatx_train_epoch(gatx_ctx* ctx, int32_t t, float* avg_loss, float* accuracy) {
  if (!ctx || t < 1) return fail(ctx, GATX_ERR_INVALID, "epoch index is 1-based");
  CK(cudaSetDevice(ctx->device));
  ctx->spans_used = 0;
  ctx->comm_spans_used = 0;
  ctx->epoch_replayed = false;
  int rc;
  if (epoch_graph_wanted(ctx)) {
    // forward + backward replayed as one CUDA graph; the optimizer (its bias correction depends on t) and the
    // loss read-back stay ordinary launches behind it
    if ((rc = ensure_buffers(ctx))) return rc;
    if (!ctx->have_params) return fail(ctx, GATX_ERR_INVALID, "parameters not initialised");
    if ((!ctx->epoch_exec || ctx->epoch_exec_gen != ctx->gen) && (rc = capture_epoch(ctx))) return rc;
    if (ctx->epoch_exec) {
      CK(cudaGraphLaunch(ctx->epoch_exec, ctx->st));
      ctx->launches += ctx->epoch_exec_launches;
      ctx->fwd_valid = true;
      ctx->last_count = ctx->epoch_exec_count;
      ctx->epoch_replayed = true;
      if ((rc = do_step(ctx, t))) return rc;
      CK(cudaGetLastError());
      if (avg_loss || accuracy) return read_loss(ctx, avg_loss, accuracy, ctx->last_count);
      return GATX_OK;
    }
  }
  {
    PhaseTimer whole(ctx, PH_EPOCH);
    if ((rc = do_forward(ctx))) return rc;
    if ((rc = do_backward(ctx))) return rc;
    if ((rc = do_step(ctx, t))) return rc;
  }
  CK(cudaGetLastError());
  if (avg_loss || accuracy) return read_loss(ctx, avg_loss, accuracy, ctx->last_count);
  return GATX_OK;
}

int gatx_set_slopes(gatx_ctx* ctx, float attn_slope, float act_slope) {
  // the streaming kernels evaluate LeakyReLU as max(x, slope * x), valid for 0 <= slope < 1
  if (!ctx || !(attn_slope >= 0.f && attn_slope < 1.f) || !(act_slope >= 0.f && act_slope < 1.f))
    return fail(ctx, GATX_ERR_INVALID, "LeakyReLU slopes must lie in [0, 1)");
  ctx->slopes = Slopes{attn_slope, act_slope};
  ctx->fwd_valid = false;  // activations of an earlier forward no longer match
  ++ctx->gen;              // kernel arguments of a captured epoch
  return GATX_OK;
}

int gatx_set_bias(gatx_ctx* ctx, int32_t on) {
  if (!ctx) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if ((on != 0) == (ctx->use_bias != 0)) return GATX_OK;
  if (ctx->have_bufs) free_bufs(ctx);  // the parameter layout changes: parameters must be set / initialised again
  ctx->use_bias = on != 0;
  ctx->fwd_valid = false;
  ++ctx->gen;
  return GATX_OK;
}

int gatx_set_bias_values(gatx_ctx* ctx, int32_t layer, const float* b) {
  if (!ctx || layer < 0 || layer >= ctx->L || !b) return fail(ctx, GATX_ERR_INVALID, "bad set_bias_values");
  if (!ctx->use_bias) return fail(ctx, GATX_ERR_INVALID, "gatx_set_bias(ctx, 1) first");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  const Layer& ly = ctx->layers[layer];
  CK(cudaMemcpyAsync(ctx->params + ly.b_off, b, sizeof(float) * (size_t)ly.F, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return GATX_OK;
}

int gatx_set_dropout(gatx_ctx* ctx, float p, uint64_t seed) {
  if (!ctx || !(p >= 0.f && p < 1.f)) return fail(ctx, GATX_ERR_INVALID, "dropout probability must lie in [0, 1)");
  ctx->p_drop = p;
  ctx->drop_seed = seed;
  ctx->drop_step = 0;
  ctx->fwd_valid = false;
  return GATX_OK;
}

int gatx_set_attn_dropout(gatx_ctx* ctx, float p, uint64_t seed) {
  if (!ctx || !(p >= 0.f && p < 1.f)) return fail(ctx, GATX_ERR_INVALID, "dropout probability must lie in [0, 1)");
  ctx->p_adrop = p;
  ctx->adrop_seed = seed;
  ctx->adrop_step = 0;
  ctx->ascale_key = -1;
  ctx->fwd_valid = false;
  return GATX_OK;
}

int gatx_set_cuda_graph(gatx_ctx* ctx, int32_t mode) {
  if (!ctx || mode < -1 || mode > 1) return fail(ctx, GATX_ERR_INVALID, "mode is -1 (auto), 0 (off) or 1 (on)");
  ctx->graph_mode = mode;
  return GATX_OK;
}
int gatx_cuda_graph_active(const gatx_ctx* ctx) { return ctx && ctx->epoch_replayed ? 1 : 0; }

int gatx_sync(gatx_ctx* ctx) {
  if (!ctx) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->st_comm) CK(cudaStreamSynchronize(ctx->st_comm));
  return GATX_OK;
}

int gatx_enable_timing(gatx_ctx* ctx, int32_t on) {
  if (!ctx) return GATX_ERR_INVALID;
  ctx->timing = on != 0;
  ctx->spans_used = 0;
  ++ctx->gen;
  for (auto& l : ctx->layers) l.kev_fwd = l.kev_bwd = false;
  return GATX_OK;
}

int gatx_get_timing(gatx_ctx* ctx, float* out_ms, int32_t n) {
  if (!ctx || !out_ms) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->st));
  for (int p = 0; p < PH_COUNT; ++p) ctx->phase_ms[p] = 0.f;
  for (size_t i = 0; i < ctx->spans_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b) == cudaSuccess) ctx->phase_ms[ctx->spans[i].phase] += ms;
  }
  for (int p = 0; p < n && p < PH_COUNT; ++p) out_ms[p] = ctx->phase_ms[p];
  return GATX_OK;
}

int gatx_timer_start(gatx_ctx* ctx) {
  if (!ctx) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->sw_a) {
    CK(cudaEventCreate(&ctx->sw_a));
    CK(cudaEventCreate(&ctx->sw_b));
  }
  CK(cudaEventRecord(ctx->sw_a, ctx->st));
  return GATX_OK;
}

int gatx_timer_stop(gatx_ctx* ctx, float* elapsed_ms) {
  if (!ctx || !elapsed_ms || !ctx->sw_a) return fail(ctx, GATX_ERR_INVALID, "timer not started");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->sw_b, ctx->st));
  CK(cudaEventSynchronize(ctx->sw_b));
  CK(cudaEventElapsedTime(elapsed_ms, ctx->sw_a, ctx->sw_b));
  return GATX_OK;
}

int gatx_get_edge_kernel_ms(gatx_ctx* ctx, int32_t layer, float* out3) {
  if (!ctx || !out3 || layer < 0 || layer >= ctx->L) return fail(ctx, GATX_ERR_INVALID, "bad layer");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->st));
  Layer& ly = ctx->layers[layer];
  out3[0] = out3[1] = out3[2] = 0.f;
  const bool stream = ly.vec && ctx->use_stream && edge_stream_supported(ly.H, ly.D);
  if (!stream) return GATX_OK;  // only the streaming kernels are individually timed
  if (ly.kev_fwd) cudaEventElapsedTime(&out3[0], ly.kev[0], ly.kev[1]);
  if (ly.kev_bwd) {
    cudaEventElapsedTime(&out3[1], ly.kev[2], ly.kev[3]);
    cudaEventElapsedTime(&out3[2], ly.kev[4], ly.kev[5]);
  }
  cudaGetLastError();
  return GATX_OK;
}

int64_t gatx_launch_count(const gatx_ctx* ctx) { return ctx ? ctx->launches : -1; }

int gatx_edge_bytes(gatx_ctx* ctx, int32_t layer, double* fwd_bytes, double* bwd_bytes) {
  if (!ctx || !ctx->have_bufs || layer < 0 || layer >= ctx->L) return fail(ctx, GATX_ERR_INVALID, "bad layer");
  const Layer& ly = ctx->layers[layer];
  const double N = ctx->n_rows, E = (double)ctx->E, F = ly.F, H = ly.H, b = 4.0;
  // algorithmic bytes of the fused passes with b-byte projected features
  if (fwd_bytes) *fwd_bytes = 4.0 * (N + 1) + E * (4.0 + b * F) + N * F * (b + 4.0) + 4.0 * H * E;
  if (bwd_bytes) *bwd_bytes = 8.0 * (N + 1) + E * (12.0 + (b + 4.0) * F) + N * F * (b + 12.0) + 16.0 * H * E;
  return GATX_OK;
}

int64_t gatx_tensor_size(gatx_ctx* ctx, int32_t which, int32_t layer) {
  if (!ctx) return -1;
  const bool per_layer = !(which == GATX_T_WO || which == GATX_T_GWO || which == GATX_T_Y || which == GATX_T_Z ||
                           which == GATX_T_PRED || (which >= GATX_T_COO_SRC && which <= GATX_T_CSC_EID));
  if (per_layer && (layer < 0 || layer >= ctx->L)) return -1;
  if (which >= GATX_T_COO_SRC && which <= GATX_T_CSC_EID) {
    if (!ctx->have_graph) return -1;
    switch (which) {
      case GATX_T_IN_DEGREE: return ctx->n_rows;
      case GATX_T_CSC_PTR: return (int64_t)ctx->N + 1;
      default: return ctx->E;
    }
  }
  if (!ctx->have_bufs) return -1;
  const Layer* ly = per_layer ? &ctx->layers[layer] : nullptr;
  const int DL = ctx->outdims[ctx->L - 1];
  switch (which) {
    case GATX_T_W: case GATX_T_GW: return (int64_t)ly->F * 2 * ly->I;
    case GATX_T_A: case GATX_T_GA: return ly->F;
    case GATX_T_B: case GATX_T_GB: return ly->b_off >= 0 ? ly->F : -1;
    case GATX_T_WO: case GATX_T_GWO: return (int64_t)ctx->C * DL;
    case GATX_T_PL: case GATX_T_GPL: return (int64_t)ctx->N * ly->F;
    case GATX_T_PR: case GATX_T_GPR: case GATX_T_HPRE: case GATX_T_GH: return (int64_t)ctx->n_rows * ly->F;
    case GATX_T_SCORE: case GATX_T_ALPHA: case GATX_T_GALPHA: case GATX_T_GE: return ctx->E * ly->H;
    case GATX_T_HOUT: return (int64_t)ctx->n_rows * ly->Fout;
    case GATX_T_Y: case GATX_T_Z: return (int64_t)ctx->n_rows * ctx->C;
    case GATX_T_PRED: return ctx->n_rows;
    default: return -1;
  }
}

// ---- checkpoint: [params | Adam m | Adam v], each n_params floats in the flat order W_0..W_{L-1} | a_0.. | W_o --------
int64_t gatx_state_size(gatx_ctx* ctx) {
  if (!ctx || ensure_buffers(ctx)) return -1;
  return 3 * ctx->n_params;
}
int gatx_get_state(gatx_ctx* ctx, float* dst, size_t bytes) {
  if (!ctx || !dst) return fail(ctx, GATX_ERR_INVALID, "bad get_state");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  const size_t nb = sizeof(float) * (size_t)ctx->n_params;
  if (bytes != 3 * nb) return fail(ctx, GATX_ERR_INVALID, "state needs %zu bytes, got %zu", 3 * nb, bytes);
  CK(cudaMemcpyAsync(dst, ctx->params, nb, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaMemcpyAsync(dst + ctx->n_params, ctx->adam_m, nb, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaMemcpyAsync(dst + 2 * ctx->n_params, ctx->adam_v, nb, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return GATX_OK;
}
int gatx_set_state(gatx_ctx* ctx, const float* src, size_t bytes) {
  if (!ctx || !src) return fail(ctx, GATX_ERR_INVALID, "bad set_state");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  const size_t nb = sizeof(float) * (size_t)ctx->n_params;
  if (bytes != 3 * nb) return fail(ctx, GATX_ERR_INVALID, "state needs %zu bytes, got %zu", 3 * nb, bytes);
  CK(cudaMemcpyAsync(ctx->params, src, nb, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(ctx->adam_m, src + ctx->n_params, nb, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(ctx->adam_v, src + 2 * ctx->n_params, nb, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->have_params = true;
  return GATX_OK;
}

int gatx_get_tensor(gatx_ctx* ctx, int32_t which, int32_t layer, void* dst, size_t bytes) {
  if (!ctx || !dst) return fail(ctx, GATX_ERR_INVALID, "bad get_tensor");
  CK(cudaSetDevice(ctx->device));
  const int64_t n = gatx_tensor_size(ctx, which, layer);
  if (n < 0) return fail(ctx, GATX_ERR_INVALID, "tensor %d/%d not available", which, layer);
  if (bytes != (size_t)n * 4) return fail(ctx, GATX_ERR_INVALID, "tensor %d needs %lld bytes, got %zu", which, (long long)n * 4, bytes);
  const void* src = nullptr;
  Layer* ly = (layer >= 0 && layer < ctx->L) ? &ctx->layers[layer] : nullptr;
  float* tmp = nullptr;
  switch (which) {
    case GATX_T_W: src = ctx->params + ly->w_off; break;
    case GATX_T_A: src = ctx->params + ly->a_off; break;
    case GATX_T_WO: src = ctx->params + ctx->wo_off; break;
    case GATX_T_GW: src = ctx->grads + ly->w_off; break;
    case GATX_T_GA: src = ctx->grads + ly->a_off; break;
    case GATX_T_B: src = ctx->params + ly->b_off; break;
    case GATX_T_GB: src = ctx->grads + ly->b_off; break;
    case GATX_T_GWO: src = ctx->grads + ctx->wo_off; break;
    case GATX_T_PL: src = ly->Pl; break;
    case GATX_T_PR: src = ly->Pr; break;
    case GATX_T_SCORE: src = ly->score; break;
    case GATX_T_ALPHA:
      CK(cudaMalloc(&tmp, (size_t)n * 4 + 4));
      LAUNCHED(launch_alpha_from_score(ly->score, ctx->coo_dst, ly->mx, ly->sinv, ctx->E, ly->H, tmp, ctx->st));
      src = tmp;
      break;
    case GATX_T_HPRE: src = ly->hpre; break;
    case GATX_T_HOUT: src = ly->Hout; break;
    case GATX_T_Y: src = ctx->y; break;
    case GATX_T_Z: src = (ctx->keep_debug || ctx->gemm_mode == GATX_GEMM_TF32_TC) ? ctx->z_dbg : nullptr; break;
    case GATX_T_GH: src = ly->gH; break;
    case GATX_T_PRED: src = ctx->pred; break;
    case GATX_T_COO_SRC: src = ctx->coo_src; break;
    case GATX_T_COO_DST: src = ctx->coo_dst; break;
    case GATX_T_IN_DEGREE: src = ctx->in_deg; break;
    case GATX_T_CSC_PTR: src = ctx->csc_ptr; break;
    case GATX_T_CSC_DST: src = ctx->csc_dst; break;
    case GATX_T_CSC_EID: src = ctx->csc_eid; break;
    case GATX_T_GPL: src = ly->gPl_dbg; break;
    case GATX_T_GPR: src = ly->gPr_dbg; break;
    case GATX_T_GALPHA: src = ly->galpha_dbg; break;
    case GATX_T_GE: src = ly->ge_dbg; break;
    default: break;
  }
  if (!src) return fail(ctx, GATX_ERR_INVALID, "tensor %d needs keep_debug=1", which);
  cudaError_t e = cudaSuccess;
  if (n && (which == GATX_T_Y || which == GATX_T_Z) && ctx->ldc != ctx->C)
    e = cudaMemcpy2DAsync(dst, sizeof(float) * ctx->C, src, sizeof(float) * ctx->ldc, sizeof(float) * ctx->C, ctx->n_rows,
                          cudaMemcpyDeviceToHost, ctx->st);
  else if (n)
    e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
  if (tmp) cudaFree(tmp);
  if (e != cudaSuccess) return fail(ctx, GATX_ERR_CUDA, "get_tensor copy: %s", cudaGetErrorString(e));
  return GATX_OK;
}

int gatx_op_gemm(int32_t mode, int32_t form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                 int64_t ldc, int32_t M, int32_t N, int64_t K) {
  // host-pointer GEMM used by the parity tests and ncu: form 0: C = A[M][K] B[N][K]^T; form 1: C = A[K][M]^T B[K][N]
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return GATX_ERR_INVALID;
  const int64_t a_rows = form == 0 ? M : K, b_rows = form == 0 ? N : K;
  float *dA = nullptr, *dB = nullptr, *dC = nullptr, *ws = nullptr;
  const size_t ws_bytes = (size_t)64 << 20;
  int rc = GATX_OK;
  cudaStream_t st = nullptr;
  if (cudaMalloc(&dA, sizeof(float) * a_rows * lda) != cudaSuccess || cudaMalloc(&dB, sizeof(float) * b_rows * ldb) != cudaSuccess ||
      cudaMalloc(&dC, sizeof(float) * (size_t)M * ldc) != cudaSuccess || cudaMalloc(&ws, ws_bytes) != cudaSuccess ||
      cudaStreamCreate(&st) != cudaSuccess)
    rc = GATX_ERR_CUDA;
  if (!rc) {
    cudaMemcpyAsync(dA, A, sizeof(float) * a_rows * lda, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dB, B, sizeof(float) * b_rows * ldb, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(dC, 0, sizeof(float) * (size_t)M * ldc, st);
    int n = -1;
    if (form == 0)
      n = mode == GATX_GEMM_TF32_TC ? launch_gemm_tc_tn(dA, lda, dB, ldb, dC, ldc, M, N, (int)K, false, st)
                                    : launch_gemm_simt(dA, lda, 1, dB, ldb, 1, dC, ldc, M, N, K, false, nullptr, 0, st);
    else
      n = mode == GATX_GEMM_TF32_TC ? launch_gemm_tc_atb(dA, lda, dB, ldb, dC, ldc, M, N, K, ws, ws_bytes, st)
                                    : launch_gemm_simt(dA, 1, lda, dB, 1, ldb, dC, ldc, M, N, K, true, ws, ws_bytes, st);
    if (n < 0) rc = GATX_ERR_UNSUPPORTED;
    if (!rc && cudaMemcpyAsync(C, dC, sizeof(float) * (size_t)M * ldc, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = GATX_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = rc ? rc : GATX_ERR_CUDA;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(ws);
  if (st) cudaStreamDestroy(st);
  return rc;
}

// ---- op-level entry points on host buffers: one kernel family at a time, for parity tests and ncu ------
// The edge ops build a throw-away one-layer context around the caller's graph so that they run exactly the kernels
// (family selection, chunking, fix-ups) the epoch runs for that (heads, outdim) shape.
namespace {
struct OpCtx {
  gatx_ctx* c = nullptr;
  ~OpCtx() { gatx_destroy(c); }
};
int op_make_ctx(OpCtx& o, int32_t N, int64_t E, const int32_t* row_ptr, const int32_t* col_idx, int32_t H, int32_t D) {
  if (N <= 0 || E < 0 || !row_ptr || H <= 0 || D <= 0) return GATX_ERR_INVALID;
  int dev = 0;
  cudaGetDevice(&dev);
  gatx_config cfg{};
  const int32_t heads[1] = {H}, outdims[1] = {D};
  cfg.num_layers = 1; cfg.heads = heads; cfg.outdims = outdims; cfg.optimizer = GATX_OPT_SGD; cfg.lr = 0.f;
  cfg.beta1 = 0.9f; cfg.beta2 = 0.999f; cfg.device = dev; cfg.gemm_mode = GATX_GEMM_FP32_SIMT; cfg.keep_debug = 1;
  cfg.rank = 0; cfg.world = 1;
  int rc = gatx_create(&o.c, &cfg);
  if (rc) return rc;
  if ((rc = gatx_set_graph_csr(o.c, N, E, row_ptr, col_idx))) return rc;
  std::vector<float> x((size_t)N * 4, 0.f);
  std::vector<int32_t> y((size_t)N, 0);
  if ((rc = gatx_set_features(o.c, x.data(), 4))) return rc;
  if ((rc = gatx_set_labels(o.c, y.data(), 1))) return rc;
  return ensure_buffers(o.c);
}
int op_forward(gatx_ctx* ctx, const float* Pl, const float* Pr, const float* a) {
  Layer& ly = ctx->layers[0];
  const size_t nf = sizeof(float) * (size_t)ctx->N * ly.F;
  CK(cudaMemcpyAsync(ly.Pl, Pl, nf, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(ly.Pr, Pr, nf, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(ctx->params + ly.a_off, a, sizeof(float) * ly.F, cudaMemcpyHostToDevice, ctx->st));
  return fwd_edge(ctx, 0, row_view(ctx, -1));
}
}  // namespace

int gatx_op_edge_fwd(int32_t N, int64_t E, const int32_t* row_ptr, const int32_t* col_idx, int32_t H, int32_t D,
                     const float* Pl, const float* Pr, const float* a, float* score, float* alpha, float* hpre,
                     float* Hout) {
  if (!Pl || !Pr || !a) return GATX_ERR_INVALID;
  OpCtx o;
  int rc = op_make_ctx(o, N, E, row_ptr, col_idx, H, D);
  if (rc) return rc;
  gatx_ctx* ctx = o.c;
  if ((rc = op_forward(ctx, Pl, Pr, a))) return rc;
  Layer& ly = ctx->layers[0];
  const size_t nf = sizeof(float) * (size_t)N * ly.F, ne = sizeof(float) * (size_t)E * H;
  if (score && E) CK(cudaMemcpyAsync(score, ly.score, ne, cudaMemcpyDeviceToHost, ctx->st));
  if (alpha && E) {
    LAUNCHED(launch_alpha_from_score(ly.score, ctx->coo_dst, ly.mx, ly.sinv, E, H, ly.alpha_dbg, ctx->st));
    CK(cudaMemcpyAsync(alpha, ly.alpha_dbg, ne, cudaMemcpyDeviceToHost, ctx->st));
  }
  if (hpre) CK(cudaMemcpyAsync(hpre, ly.hpre, nf, cudaMemcpyDeviceToHost, ctx->st));
  if (Hout) CK(cudaMemcpyAsync(Hout, ly.Hfull, nf, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaGetLastError());
  return GATX_OK;
}

int gatx_op_edge_bwd(int32_t N, int64_t E, const int32_t* row_ptr, const int32_t* col_idx, int32_t H, int32_t D,
                     const float* Pl, const float* Pr, const float* a, const float* gHout, float* g_pre, float* gPl,
                     float* gPr, float* ga, float* ge) {
  if (!Pl || !Pr || !a || !gHout) return GATX_ERR_INVALID;
  OpCtx o;
  int rc = op_make_ctx(o, N, E, row_ptr, col_idx, H, D);
  if (rc) return rc;
  gatx_ctx* ctx = o.c;
  if ((rc = op_forward(ctx, Pl, Pr, a))) return rc;
  Layer& ly = ctx->layers[0];
  const size_t nf = sizeof(float) * (size_t)N * ly.F, ne = sizeof(float) * (size_t)E * H;
  CK(cudaMemcpyAsync(ly.gH, gHout, nf, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemsetAsync(ctx->grads + ly.a_off, 0, sizeof(float) * ly.F, ctx->st));
  const RowView all = row_view(ctx, -1);
  const bool stream = ly.vec && ctx->use_stream && edge_stream_supported(ly.H, ly.D);
  if (stream) {
    if ((rc = bwd_edge(ctx, 0, all, 1))) return rc;
    if ((rc = bwd_edge(ctx, 0, all, 2))) return rc;
  } else if ((rc = bwd_edge(ctx, 0, all, 3))) {
    return rc;
  }
  if (g_pre) CK(cudaMemcpyAsync(g_pre, ly.gH, nf, cudaMemcpyDeviceToHost, ctx->st));
  if (gPl) CK(cudaMemcpyAsync(gPl, ctx->gPl, nf, cudaMemcpyDeviceToHost, ctx->st));
  if (gPr) CK(cudaMemcpyAsync(gPr, ctx->gPr, nf, cudaMemcpyDeviceToHost, ctx->st));
  if (ga) CK(cudaMemcpyAsync(ga, ctx->grads + ly.a_off, sizeof(float) * ly.F, cudaMemcpyDeviceToHost, ctx->st));
  if (ge && E) {
    if (ly.vec) LAUNCHED(launch_unpack_rec(ctx->rec, E, ly.H, ly.D, ly.alpha_dbg, ly.ge_dbg, ctx->st));
    else LAUNCHED(launch_unpack_rec_generic(ctx->rec, E, ly.H, ly.alpha_dbg, ly.ge_dbg, ctx->st));
    CK(cudaMemcpyAsync(ge, ly.ge_dbg, ne, cudaMemcpyDeviceToHost, ctx->st));
  }
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaGetLastError());
  return GATX_OK;
}

int gatx_op_softmax_ce(int32_t N, int32_t C, const float* z, const int32_t* labels, const uint8_t* mask, float* y,
                       float* dz, int32_t* pred, double* loss_sum, int64_t* correct) {
  if (N <= 0 || C <= 0 || !z || !labels) return GATX_ERR_INVALID;
  for (int i = 0; i < N; ++i)
    if (labels[i] < 0 || labels[i] >= C) return GATX_ERR_INVALID;
  const int ldc = (C + 3) / 4 * 4;
  float *dzin = nullptr, *dy = nullptr, *ddz = nullptr;
  int *dl = nullptr, *dp = nullptr, *cp = nullptr;
  unsigned char* dm = nullptr;
  double *lp = nullptr, *ls = nullptr;
  long long* cc = nullptr;
  cudaStream_t st = nullptr;
  int rc = GATX_OK, n_part = 0;
  const size_t nz = sizeof(float) * (size_t)N * ldc;
  if (cudaMalloc(&dzin, nz) != cudaSuccess || cudaMalloc(&dy, nz) != cudaSuccess || cudaMalloc(&ddz, nz) != cudaSuccess ||
      cudaMalloc(&dl, sizeof(int) * N) != cudaSuccess || cudaMalloc(&dp, sizeof(int) * N) != cudaSuccess ||
      cudaMalloc(&cp, sizeof(int) * kHeadBlocks) != cudaSuccess || cudaMalloc(&lp, sizeof(double) * kHeadBlocks) != cudaSuccess ||
      cudaMalloc(&ls, sizeof(double)) != cudaSuccess || cudaMalloc(&cc, sizeof(long long)) != cudaSuccess ||
      (mask && cudaMalloc(&dm, N) != cudaSuccess) || cudaStreamCreate(&st) != cudaSuccess)
    rc = GATX_ERR_CUDA;
  if (!rc) {
    cudaMemsetAsync(dzin, 0, nz, st);
    cudaMemcpy2DAsync(dzin, sizeof(float) * ldc, z, sizeof(float) * C, sizeof(float) * C, N, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dl, labels, sizeof(int) * N, cudaMemcpyHostToDevice, st);
    if (mask) cudaMemcpyAsync(dm, mask, N, cudaMemcpyHostToDevice, st);
    int n1 = launch_softmax_ce(dzin, dl, N, C, ldc, dy, ddz, dp, lp, cp, &n_part, dm, st);
    int n2 = n1 < 0 ? -1 : launch_loss_finalize(lp, cp, n_part, ls, cc, st);
    if (n1 < 0 || n2 < 0) rc = GATX_ERR_UNSUPPORTED;
    if (!rc) {
      if (y) cudaMemcpy2DAsync(y, sizeof(float) * C, dy, sizeof(float) * ldc, sizeof(float) * C, N, cudaMemcpyDeviceToHost, st);
      if (dz) cudaMemcpy2DAsync(dz, sizeof(float) * C, ddz, sizeof(float) * ldc, sizeof(float) * C, N, cudaMemcpyDeviceToHost, st);
      if (pred) cudaMemcpyAsync(pred, dp, sizeof(int) * N, cudaMemcpyDeviceToHost, st);
      long long ch = 0;
      if (loss_sum) cudaMemcpyAsync(loss_sum, ls, sizeof(double), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(&ch, cc, sizeof(long long), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = GATX_ERR_CUDA;
      if (correct) *correct = (int64_t)ch;
    }
  }
  cudaFree(dzin); cudaFree(dy); cudaFree(ddz); cudaFree(dl); cudaFree(dp); cudaFree(cp); cudaFree(lp); cudaFree(ls);
  cudaFree(cc); cudaFree(dm);
  if (st) cudaStreamDestroy(st);
  return rc;
}

int gatx_op_optimizer(int64_t n, const int64_t* group_end3, int32_t optimizer, int32_t clip, float lr, float beta1,
                      float beta2, int32_t t, float* params, float* grads, float* adam_m, float* adam_v) {
  if (n <= 0 || !group_end3 || !params || !grads || t < 1) return GATX_ERR_INVALID;
  if (optimizer == GATX_OPT_ADAM && (!adam_m || !adam_v)) return GATX_ERR_INVALID;
  if (!(0 <= group_end3[0] && group_end3[0] <= group_end3[1] && group_end3[1] <= group_end3[2] && group_end3[2] == n))
    return GATX_ERR_INVALID;
  float *dp = nullptr, *dg = nullptr, *dm = nullptr, *dv = nullptr, *np_ = nullptr;
  cudaStream_t st = nullptr;
  int rc = GATX_OK;
  const size_t nb = sizeof(float) * (size_t)n;
  if (cudaMalloc(&dp, nb) != cudaSuccess || cudaMalloc(&dg, nb) != cudaSuccess || cudaMalloc(&dm, nb) != cudaSuccess ||
      cudaMalloc(&dv, nb) != cudaSuccess || cudaMalloc(&np_, sizeof(float) * 6 * kOptimBlocks) != cudaSuccess ||
      cudaStreamCreate(&st) != cudaSuccess)
    rc = GATX_ERR_CUDA;
  if (!rc) {
    cudaMemcpyAsync(dp, params, nb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dg, grads, nb, cudaMemcpyHostToDevice, st);
    if (adam_m) cudaMemcpyAsync(dm, adam_m, nb, cudaMemcpyHostToDevice, st); else cudaMemsetAsync(dm, 0, nb, st);
    if (adam_v) cudaMemcpyAsync(dv, adam_v, nb, cudaMemcpyHostToDevice, st); else cudaMemsetAsync(dv, 0, nb, st);
    OptimGroups grp{};
    grp.begin[0] = 0; grp.end[0] = group_end3[0];
    grp.begin[1] = group_end3[0]; grp.end[1] = group_end3[1];
    grp.begin[2] = group_end3[1]; grp.end[2] = group_end3[2];
    if (launch_optimizer(dp, dg, dm, dv, n, grp, clip != 0, optimizer, lr, beta1, beta2, t, np_, st) < 0) rc = GATX_ERR_UNSUPPORTED;
    cudaMemcpyAsync(params, dp, nb, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(grads, dg, nb, cudaMemcpyDeviceToHost, st);  // zeroed by the update (EB:1631-1633)
    if (adam_m) cudaMemcpyAsync(adam_m, dm, nb, cudaMemcpyDeviceToHost, st);
    if (adam_v) cudaMemcpyAsync(adam_v, dv, nb, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = rc ? rc : GATX_ERR_CUDA;
  }
  cudaFree(dp); cudaFree(dg); cudaFree(dm); cudaFree(dv); cudaFree(np_);
  if (st) cudaStreamDestroy(st);
  return rc;
}

int gatx_comm_unique_id(void* out128) {
  if (!out128 || !g_nccl.load()) return GATX_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return GATX_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  memcpy(out128, &id, 128);
  return GATX_OK;
}

int gatx_comm_init(gatx_ctx* ctx, const void* id128) {
  if (!ctx || !id128) return GATX_ERR_INVALID;
  if (!g_nccl.load()) return fail(ctx, GATX_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror());
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NK(g_nccl.CommInitRank(&ctx->comm, ctx->world, id, ctx->rank));
  return GATX_OK;
}

// ---- NVLink peer-memory halo exchange: handle exchange ------------------------------------------------------------
namespace {
constexpr int kPeerMaxBufs = 24;
struct PeerInfo {  // what one rank publishes; sizeof <= GATX_PEER_INFO_BYTES
  int32_t magic, pid, device, n_bufs;  // buffers: P_l of layer 0..L-1, the gP_l scratch, the barrier flags
  int64_t n_floats[kPeerMaxBufs];
  uint64_t raw[kPeerMaxBufs];
  cudaIpcMemHandle_t handle[kPeerMaxBufs];
};
static_assert(sizeof(PeerInfo) <= GATX_PEER_INFO_BYTES, "PeerInfo must fit the public blob");
constexpr int32_t kPeerMagic = 0x47585033;  // "GXP3"
}  // namespace

int gatx_peer_export(gatx_ctx* ctx, void* out, size_t bytes) {
  if (!ctx || !out || bytes < GATX_PEER_INFO_BYTES) return fail(ctx, GATX_ERR_INVALID, "bad peer_export");
  if (ctx->world < 2 || ctx->world > kMaxPeers) return fail(ctx, GATX_ERR_INVALID, "peer exchange needs 2..%d ranks", kMaxPeers);
  if (ctx->L + 3 > kPeerMaxBufs) return fail(ctx, GATX_ERR_UNSUPPORTED, "too many layers for the peer blob");
  if (!ctx->my_ref) return fail(ctx, GATX_ERR_INVALID, "the graph was set for a single rank");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_buffers(ctx);
  if (rc) return rc;
  CK(cudaMemsetAsync(ctx->halo_flags, 0, sizeof(uint32_t) * kMaxPeers, ctx->st));  // a new exchange starts at barrier 0
  ctx->barrier_seq = 0;
  CK(cudaStreamSynchronize(ctx->st));  // the flags are zero before any peer can see the handle
  PeerInfo info{};
  info.magic = kPeerMagic;
  info.pid = (int32_t)getpid();
  info.device = ctx->device;
  info.n_bufs = ctx->L + 3;
  int64_t Fmax = 0;
  for (int l = 0; l < ctx->L; ++l) Fmax = ctx->layers[l].F > Fmax ? ctx->layers[l].F : Fmax;
  for (int b = 0; b <= ctx->L + 2; ++b) {
    // buffers: P_l of layer 0..L-1, the staging buffer of the backward exchange, the barrier flags, the gP_l scratch
    void* ptr = b < ctx->L ? (void*)ctx->layers[b].Pl
                           : (b == ctx->L ? (void*)ctx->stage : (b == ctx->L + 1 ? (void*)ctx->halo_flags : (void*)ctx->gPl));
    info.n_floats[b] = b < ctx->L ? (int64_t)ctx->N * ctx->layers[b].F
                                  : (b == ctx->L ? (int64_t)ctx->world * ctx->n_rows * Fmax
                                                 : (b == ctx->L + 1 ? (int64_t)kMaxPeers : (int64_t)ctx->N * Fmax));
    info.raw[b] = (uint64_t)(uintptr_t)ptr;
    CK(cudaIpcGetMemHandle(&info.handle[b], ptr));
  }
  memset(out, 0, GATX_PEER_INFO_BYTES);
  memcpy(out, &info, sizeof info);
  return GATX_OK;
}

int gatx_peer_import(gatx_ctx* ctx, const void* all, size_t bytes) {
  if (!ctx || !all) return fail(ctx, GATX_ERR_INVALID, "bad peer_import");
  if (ctx->world < 2 || ctx->world > kMaxPeers || bytes != (size_t)ctx->world * GATX_PEER_INFO_BYTES)
    return fail(ctx, GATX_ERR_INVALID, "peer_import needs world * GATX_PEER_INFO_BYTES bytes in rank order");
  if (!ctx->have_bufs || !ctx->ref_mask) return fail(ctx, GATX_ERR_INVALID, "call gatx_peer_export first");
  if (!ctx->comm) return fail(ctx, GATX_ERR_INVALID, "gatx_comm_init must come first (barriers)");
  CK(cudaSetDevice(ctx->device));
  ctx->peer_Pl.assign(ctx->L, PeerPtrs{});
  ctx->peer_stage = PeerPtrs{};
  ctx->peer_gPl = PeerPtrs{};
  ctx->peer_flags = PeerFlags{};
  ctx->peer_flags.p[ctx->rank] = ctx->halo_flags;
  int64_t Fmax = 0;
  for (int l = 0; l < ctx->L; ++l) Fmax = ctx->layers[l].F > Fmax ? ctx->layers[l].F : Fmax;
  for (int p = 0; p < ctx->world; ++p) {
    PeerInfo info;
    memcpy(&info, (const char*)all + (size_t)p * GATX_PEER_INFO_BYTES, sizeof info);
    if (info.magic != kPeerMagic || info.n_bufs != ctx->L + 3)
      return fail(ctx, GATX_ERR_INVALID, "peer blob of rank %d does not match this model", p);
    for (int b = 0; b <= ctx->L; ++b)
      if (info.n_floats[b] != (b < ctx->L ? (int64_t)ctx->N * ctx->layers[b].F
                                          : (int64_t)ctx->world * (ctx->bounds[p + 1] - ctx->bounds[p]) * Fmax))
        return fail(ctx, GATX_ERR_INVALID, "peer blob of rank %d: buffer %d has another size", p, b);
    if (p == ctx->rank) continue;
    const bool same_process = info.pid == (int32_t)getpid();
    if (same_process) {
      // contexts driven by threads of one process (train_gatx --gpus N): plain peer access on the raw pointers
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, ctx->device, info.device));
      if (!can) return fail(ctx, GATX_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, info.device);
      cudaError_t e = cudaDeviceEnablePeerAccess(info.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(ctx, GATX_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    }
    for (int b = 0; b <= ctx->L + 2; ++b) {
      void* q = nullptr;
      if (same_process) {
        q = (void*)(uintptr_t)info.raw[b];
      } else {
        CK(cudaIpcOpenMemHandle(&q, info.handle[b], cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened.push_back(q);
      }
      if (b < ctx->L) ctx->peer_Pl[b].p[p] = (float*)q;
      else if (b == ctx->L) ctx->peer_stage.p[p] = (float*)q;
      else if (b == ctx->L + 1) ctx->peer_flags.p[p] = (uint32_t*)q;
      else ctx->peer_gPl.p[p] = (float*)q;
    }
  }
  ctx->peers_ready = getenv("GATX_NO_P2P") == nullptr;
  return GATX_OK;
}

int gatx_peer_disable(gatx_ctx* ctx) {
  if (!ctx) return GATX_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  if (ctx->st_comm) cudaStreamSynchronize(ctx->st_comm);
  for (void* q : ctx->ipc_opened) cudaIpcCloseMemHandle(q);
  ctx->ipc_opened.clear();
  ctx->peers_ready = false;
  return GATX_OK;
}

int gatx_halo_stats(gatx_ctx* ctx, double* out4) {
  if (!ctx || !out4) return GATX_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->st_comm) CK(cudaStreamSynchronize(ctx->st_comm));
  out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
  double lane_ms[2][1] = {};
  for (size_t i = 0; i < ctx->comm_spans_used; ++i) {
    float ms = 0.f;
    const auto& s = ctx->comm_spans[i];
    if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) continue;
    out4[2 * s.dir] += s.bytes;
    lane_ms[s.dir][0] += ms;
  }
  for (int d = 0; d < 2; ++d) out4[2 * d + 1] = lane_ms[d][0];
  cudaGetLastError();
  return GATX_OK;
}

int64_t gatx_halo_rows(const gatx_ctx* ctx) { return ctx ? ctx->halo_rows : -1; }
int gatx_halo_active(const gatx_ctx* ctx) { return ctx && ctx->peers_ready ? 1 : 0; }

}  // extern "C"

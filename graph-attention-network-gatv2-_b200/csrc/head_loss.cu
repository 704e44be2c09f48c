// Classifier head, softmax, cross-entropy, argmax and the output gradients in one pass over
// the last layer's output.  Replaces gatv2_output_kernel + softmax (EB:463-511, EB:132-141),
// compute_loss_accuracy_kernel (EB:514-537), the two thrust reductions (EB:539-550) and the
// dz / dH parts of compute_output_gradients (EB:553-608).  gW_o = dz^T H_L (EB:576-581, global
// atomics in the reference) is a deterministic split-K GEMM on the dz this kernel writes.
#include "common.cuh"

namespace gatx {

constexpr int HT = 32;  // nodes per tile

__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ HL, const float* __restrict__ Wo, const int* __restrict__ labels, int N,
            int C, int DL, int ldc, float* __restrict__ y, float* __restrict__ dz, float* __restrict__ z_dbg,
            int* __restrict__ pred, float* __restrict__ gH, double* __restrict__ loss_partials,
            int* __restrict__ correct_partials, const unsigned char* __restrict__ mask) {
  extern __shared__ __align__(16) float smem[];
  const int ldw = DL + 1, ldy = C + 1;
  float* Wo_s = smem;               // [C][DL+1]
  float* Ht = Wo_s + C * ldw;       // [HT][DL+1]
  float* ys = Ht + HT * ldw;        // [HT][C+1]  logits, then probabilities, then dz
  int* lab = reinterpret_cast<int*>(ys + HT * ldy);  // [HT]
  int* msk = lab + HT;                                // [HT] 1 = node counts (extension)
  const int tid = threadIdx.x;
  for (int i = tid; i < C * DL; i += blockDim.x) Wo_s[(i / DL) * ldw + (i % DL)] = __ldg(Wo + i);
  double loss_acc = 0.0;
  int correct_acc = 0;
  const int n_tiles = (N + HT - 1) / HT;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * HT;
    const int nn = N - n0 < HT ? N - n0 : HT;
    __syncthreads();  // previous tile fully consumed (also orders the Wo_s fill)
    for (int i = tid; i < nn * DL; i += blockDim.x) Ht[(i / DL) * ldw + (i % DL)] = __ldg(HL + (int64_t)n0 * DL + i);
    if (tid < nn) {
      lab[tid] = __ldg(labels + n0 + tid);
      msk[tid] = mask ? (mask[n0 + tid] != 0) : 1;
    }
    __syncthreads();
    // z = W_o h  (EB:493-500)
    for (int p = tid; p < nn * C; p += blockDim.x) {
      const int n = p / C, c = p % C;
      const float* w = Wo_s + c * ldw;
      const float* h = Ht + n * ldw;
      float acc = 0.f;
      for (int d = 0; d < DL; ++d) acc = fmaf(w[d], h[d], acc);
      ys[n * ldy + c] = acc;
      if (z_dbg) z_dbg[(int64_t)(n0 + n) * ldc + c] = acc;
    }
    __syncthreads();
    if (tid < nn) {
      float* row = ys + tid * ldy;
      float m = row[0];
      for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
      float sum = 0.f;
      for (int c = 0; c < C; ++c) {
        row[c] = expf(row[c] - m);  // EB:137
        sum += row[c];
      }
      const double den = (double)sum + 1e-8;  // EB:140 (double literal in the reference)
      float best = 0.f;
      int arg = 0;
      const int l = lab[tid];
      for (int c = 0; c < C; ++c) {
        const float p = (float)((double)row[c] / den);
        row[c] = p;
        if (c == 0 || p > best) {  // first maximum wins (EB:530-535)
          best = p;
          arg = c;
        }
      }
      if (msk[tid]) {
        loss_acc += (double)(-logf(fmaxf(row[l], 1e-12f)));  // EB:527
        correct_acc += (arg == l);
      }
      pred[n0 + tid] = arg;
    }
    __syncthreads();
    // y out, dz = y - onehot (EB:572) kept in smem and written for the gW_o GEMM
    for (int p = tid; p < nn * C; p += blockDim.x) {
      const int n = p / C, c = p % C;
      const float prob = ys[n * ldy + c];
      const float d = msk[n] ? prob - (c == lab[n] ? 1.0f : 0.0f) : 0.f;
      y[(int64_t)(n0 + n) * ldc + c] = prob;
      dz[(int64_t)(n0 + n) * ldc + c] = d;
      ys[n * ldy + c] = d;
    }
    __syncthreads();
    // dL/dH_L = W_o^T dz  (EB:590-594); the LReLU derivative and the 1/H factor (EB:597-602)
    // are applied by the edge backward / head-broadcast kernels
    for (int p = tid; p < nn * DL; p += blockDim.x) {
      const int n = p / DL, d = p % DL;
      const float* dzr = ys + n * ldy;
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(Wo_s[c * ldw + d], dzr[c], acc);
      gH[(int64_t)n0 * DL + p] = acc;
    }
  }
  // block reduction of the loss / correct counters held by threads 0..31 (warp 0)
  if (tid < 32) {
    for (int off = 16; off > 0; off >>= 1) {
      loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
      correct_acc += __shfl_down_sync(0xffffffffu, correct_acc, off);
    }
    if (tid == 0) {
      loss_partials[blockIdx.x] = loss_acc;
      correct_partials[blockIdx.x] = correct_acc;
    }
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ loss_partials, const int* __restrict__ correct_partials,
                                     int n_partials, double* __restrict__ loss_sum, long long* __restrict__ correct) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    long long c = 0;
    for (int i = 0; i < n_partials; ++i) {
      s += loss_partials[i];
      c += correct_partials[i];
    }
    *loss_sum = s;
    *correct = c;
  }
}

static size_t head_smem_bytes(int C, int DL) {
  return sizeof(float) * ((size_t)C * (DL + 1) + (size_t)HT * (DL + 1) + (size_t)HT * (C + 1)) + sizeof(int) * 2 * HT;
}

int launch_head(const float* HL, const float* Wo, const int* labels, int N, int C, int DL, int ldc, float* y, float* dz,
                float* z_dbg, int* pred, float* gH, double* loss_partials, int* correct_partials, int* n_partials,
                const unsigned char* mask, cudaStream_t st) {
  const size_t smem = head_smem_bytes(C, DL);
  if (smem > 200 * 1024) return -1;
  if (smem > 48 * 1024) cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks = (N + HT - 1) / HT;
  if (blocks > kHeadBlocks) blocks = kHeadBlocks;
  if (blocks < 1) blocks = 1;
  head_kernel<<<blocks, 256, smem, st>>>(HL, Wo, labels, N, C, DL, ldc, y, dz, z_dbg, pred, gH, loss_partials,
                                         correct_partials, mask);
  *n_partials = blocks;
  return 1;
}

// ---- tensor-core path: logits come from the tcgen05 GEMM, this kernel does softmax / CE / argmax / dz ----
constexpr int ST = 128;  // nodes per tile, one thread per node

__global__ void __launch_bounds__(ST)
softmax_ce_kernel(const float* __restrict__ z, const int* __restrict__ labels, int N, int C, int ldc,
                  float* __restrict__ y, float* __restrict__ dz, int* __restrict__ pred,
                  double* __restrict__ loss_partials, int* __restrict__ correct_partials,
                  const unsigned char* __restrict__ mask) {
  extern __shared__ __align__(16) float smem[];
  const int lds = ldc + 1;  // odd pitch: a thread walking its own row is bank-conflict free
  float* zt = smem;             // [ST][ldc+1] logits -> probabilities
  float* dt = smem + ST * lds;  // [ST][ldc+1] dz
  __shared__ double red_l[ST / 32];
  __shared__ int red_c[ST / 32];
  const int tid = threadIdx.x;
  double loss_acc = 0.0;
  int correct_acc = 0;
  const int n_tiles = (N + ST - 1) / ST;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * ST;
    const int nn = N - n0 < ST ? N - n0 : ST;
    __syncthreads();
    for (int i = tid; i < nn * ldc; i += ST) zt[(i / ldc) * lds + (i % ldc)] = __ldg(z + (int64_t)n0 * ldc + i);
    __syncthreads();
    if (tid < nn) {
      float* row = zt + tid * lds;
      float* drow = dt + tid * lds;
      float m = row[0];
      for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
      float sum = 0.f;
      for (int c = 0; c < C; ++c) {
        row[c] = expf(row[c] - m);  // EB:137
        sum += row[c];
      }
      const double den = (double)sum + 1e-8;  // EB:140
      const int l = __ldg(labels + n0 + tid);
      const bool counts = mask ? (mask[n0 + tid] != 0) : true;  // extension: masked-out nodes get dz = 0
      float best = 0.f;
      int arg = 0;
      for (int c = 0; c < C; ++c) {
        const float p = (float)((double)row[c] / den);
        row[c] = p;
        drow[c] = counts ? p - (c == l ? 1.0f : 0.0f) : 0.f;  // EB:572
        if (c == 0 || p > best) {              // EB:530-535
          best = p;
          arg = c;
        }
      }
      for (int c = C; c < ldc; ++c) row[c] = drow[c] = 0.f;
      if (counts) {
        loss_acc += (double)(-logf(fmaxf(row[l], 1e-12f)));  // EB:527
        correct_acc += (arg == l);
      }
      pred[n0 + tid] = arg;
    }
    __syncthreads();
    for (int i = tid; i < nn * ldc; i += ST) {
      y[(int64_t)n0 * ldc + i] = zt[(i / ldc) * lds + (i % ldc)];
      dz[(int64_t)n0 * ldc + i] = dt[(i / ldc) * lds + (i % ldc)];
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
    correct_acc += __shfl_down_sync(0xffffffffu, correct_acc, off);
  }
  if ((tid & 31) == 0) {
    red_l[tid >> 5] = loss_acc;
    red_c[tid >> 5] = correct_acc;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    int c = 0;
    for (int w = 0; w < ST / 32; ++w) {
      s += red_l[w];
      c += red_c[w];
    }
    loss_partials[blockIdx.x] = s;
    correct_partials[blockIdx.x] = c;
  }
}

int launch_softmax_ce(const float* z, const int* labels, int N, int C, int ldc, float* y, float* dz, int* pred,
                      double* loss_partials, int* correct_partials, int* n_partials, const unsigned char* mask,
                      cudaStream_t st) {
  const size_t smem = sizeof(float) * 2 * ST * (ldc + 1);
  if (smem > 200 * 1024) return -1;
  if (smem > 48 * 1024) cudaFuncSetAttribute(softmax_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks = (N + ST - 1) / ST;
  if (blocks > kHeadBlocks) blocks = kHeadBlocks;
  if (blocks < 1) blocks = 1;
  softmax_ce_kernel<<<blocks, ST, smem, st>>>(z, labels, N, C, ldc, y, dz, pred, loss_partials, correct_partials, mask);
  *n_partials = blocks;
  return 1;
}

// WoT[d][c] = Wo[c][d], row pitch ldc, zero padded (K-major B operand of gH = dz Wo)
__global__ void transpose_wo_kernel(const float* __restrict__ Wo, int C, int DL, int ldc, float* __restrict__ WoT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= DL * ldc) return;
  const int d = i / ldc, c = i % ldc;
  WoT[i] = c < C ? Wo[c * DL + d] : 0.f;
}
int launch_transpose_wo(const float* Wo, int C, int DL, int ldc, float* WoT, cudaStream_t st) {
  transpose_wo_kernel<<<(DL * ldc + 255) / 256, 256, 0, st>>>(Wo, C, DL, ldc, WoT);
  return 1;
}

int launch_loss_finalize(const double* loss_partials, const int* correct_partials, int n_partials,
                         double* loss_sum, long long* correct, cudaStream_t st) {
  loss_finalize_kernel<<<1, 32, 0, st>>>(loss_partials, correct_partials, n_partials, loss_sum, correct);
  return 1;
}

}  // namespace gatx

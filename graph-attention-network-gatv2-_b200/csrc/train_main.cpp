// train_gatx -- drop-in command line for the reference's train_edge / train_node binaries.
//
// Same dataset files, same flags, same stdout lines as GATv2_edge_based.cu:main (EB:927-1646); the epoch
// itself runs in libgatx through the C ABI (include/gatx.h).  Host C++ only: no CUDA calls here.
//   files  : <data-root>/<dataset>/{features,row_ptr,col_idx,labels}.txt          (EB:24-64, EB:1050-1100)
//   flags  : --num-layers --heads --outdims --epochs --optimizer --beta1 --beta2 --lr --clip --dataset
//            --data-root, env DATA_ROOT; unknown flags are ignored like the reference (EB:942-1077)
//   stdout : configuration block (EB:1024-1040), dataset lines (EB:1076-1111), per epoch
//            "\nEpoch %d\n", "\nAvg Loss: %f, Accuracy: %.2f%%\n", " total time: <ms> ms" (EB:1372, 547, 1641)
// Additions that do not change the defaults: --seed S, --gemm tf32|fp32|3xtf32, --gpus N (destination-row partition,
// one host thread + one context per GPU), --load-weights DIR / --dump-weights DIR (W.bin a.bin Wo.bin, raw fp32
// in the reference's layouts), --save-checkpoint FILE / --resume FILE (parameters + Adam moments + epoch counter),
// --quiet-epochs k (print only every k-th epoch), --no-cache (do not read/write the binary dataset cache
// <dataset>/.gatx_cache.bin, keyed by the size and mtime of the four text files), --load-only (load, report, exit),
// --split (README.md:134's announced train/val/test split: <dataset>/split.txt holds one token per node, 0 = train,
// 1 = validation, 2 = test; training loss / accuracy / gradients use the train nodes only, every printed epoch adds a
// "Val Loss: %f, Val Accuracy: %.2f%%" line and the run ends with "Test Loss: ..."), --eval-only (no training: one
// evaluation forward with the loaded weights / checkpoint, prints the "Avg Loss" line over all nodes or per split),
// --check-replicas (--gpus N: after training, compare parameters + optimizer state of all ranks bit for bit).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gatx.h"

namespace {

struct Mapped {
  const char* p = nullptr;
  size_t n = 0;
  bool open(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); return false; }
    n = (size_t)st.st_size;
    if (n) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { ::close(fd); return false; }
      p = (const char*)m;
    }
    ::close(fd);
    return true;
  }
  ~Mapped() { if (p) munmap((void*)p, n); }
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// One token -> float with `iss >> val` / strtof semantics.  Fast path: std::from_chars (correctly rounded, like strtof,
// so the bits are the same); anything it does not consume completely (leading '+', "inf", hex floats ...) goes through strtof.
inline bool parse_float(const char* t, const char* e, float& v) {
  const auto r = std::from_chars(t, e, v);
  if (r.ec == std::errc() && r.ptr == e) return true;
  char buf[64];
  const size_t n = (size_t)(e - t) < sizeof buf - 1 ? (size_t)(e - t) : sizeof buf - 1;
  memcpy(buf, t, n);
  buf[n] = 0;
  char* e2 = nullptr;
  v = strtof(buf, &e2);
  return e2 != buf;
}

// Parses the tokens of one line; stores at most `cap` values at dst and returns the number of tokens the reference's
// `while (iss >> val)` would have read (it stops at the first token that is not a number).
inline int parse_line(const char* q, const char* eol, float* dst, int cap) {
  int count = 0;
  while (q < eol) {
    while (q < eol && is_space(*q)) ++q;
    if (q >= eol) break;
    const char* t = q;
    while (q < eol && !is_space(*q)) ++q;
    float v;
    if (!parse_float(t, q, v)) break;
    if (count < cap) dst[count] = v;
    ++count;
  }
  return count;
}

int loader_threads(size_t bytes) {
  const char* e = getenv("GATX_LOADER_THREADS");
  int t = e ? atoi(e) : (int)std::thread::hardware_concurrency();
  if (t < 1) t = 1;
  if (t > 32) t = 32;
  const size_t by_size = bytes / (4u << 20) + 1;  // at least 4 MB of text per thread
  return (size_t)t < by_size ? t : (int)by_size;
}

// Line-oriented like the reference's load_features (EB:24-51): N = number of lines, I = tokens of the first
// line, every line must have the same count ("Inconsistent input_dim on line k", exit 1).  The file is cut at line
// boundaries into one segment per host thread: pass 1 counts the lines of every segment, pass 2 parses each line
// straight into its row of the [N][I] matrix (a products-size features.txt is 2.4 GB of text).
bool load_features(const std::string& path, std::vector<float>& out, int& n_nodes, int& in_dim) {
  Mapped f;
  n_nodes = 0;
  in_dim = 0;
  if (!f.open(path)) return true;  // the reference silently reads nothing from a missing file
  const char *base = f.p, *end = f.p + f.n;
  if (f.n == 0) return true;
  const int T = loader_threads(f.n);
  std::vector<const char*> cut(T + 1, end);
  cut[0] = base;
  for (int t = 1; t < T; ++t) {
    const char* guess = base + f.n / T * t;
    if (guess < cut[t - 1]) guess = cut[t - 1];
    const char* nl = guess < end ? (const char*)memchr(guess, '\n', (size_t)(end - guess)) : nullptr;
    cut[t] = nl ? nl + 1 : end;
  }
  // tokens of the first line fix I (EB:38-41)
  {
    const char* eol = (const char*)memchr(base, '\n', f.n);
    in_dim = parse_line(base, eol ? eol : end, nullptr, 0);
  }
  std::vector<int64_t> lines(T, 0);
  auto for_segments = [&](auto&& fn) {
    if (T == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(fn, t);
    for (auto& x : th) x.join();
  };
  for_segments([&](int t) {
    int64_t n = 0;
    const char* p = cut[t];
    while (p < cut[t + 1]) {
      const char* eol = (const char*)memchr(p, '\n', (size_t)(cut[t + 1] - p));
      ++n;
      p = eol ? eol + 1 : cut[t + 1];
    }
    lines[t] = n;
  });
  std::vector<int64_t> first(T + 1, 0);
  for (int t = 0; t < T; ++t) first[t + 1] = first[t] + lines[t];
  const int64_t N = first[T];
  out.resize((size_t)N * (size_t)in_dim);
  std::vector<int64_t> bad(T, -1);  // first line of the segment whose token count differs from I
  for_segments([&](int t) {
    const char* p = cut[t];
    int64_t row = first[t];
    while (p < cut[t + 1]) {
      const char* eol = (const char*)memchr(p, '\n', (size_t)(cut[t + 1] - p));
      if (!eol) eol = cut[t + 1];
      const int count = parse_line(p, eol, in_dim ? out.data() + (size_t)row * in_dim : nullptr, in_dim);
      if (count != in_dim && bad[t] < 0) bad[t] = row;
      ++row;
      p = eol < cut[t + 1] ? eol + 1 : cut[t + 1];
    }
  });
  for (int t = 0; t < T; ++t)
    if (bad[t] >= 0) {
      // in_dim == 0 (an empty first line) adopts the next line's count in the reference; not worth a fast path
      std::cerr << "Inconsistent input_dim on line " << bad[t] << std::endl;
      exit(1);
    }
  n_nodes = (int)N;
  return true;
}

// Whitespace-separated integers (EB:53-64), parsed by one host thread per segment of the file; `file >> int` stops at
// the first token that is not an integer, so everything after such a token is dropped.
void load_int_array(const std::string& path, std::vector<int>& out) {
  Mapped f;
  if (!f.open(path) || f.n == 0) return;
  const char *base = f.p, *end = f.p + f.n;
  const int T = loader_threads(f.n);
  std::vector<const char*> cut(T + 1, end);
  cut[0] = base;
  for (int t = 1; t < T; ++t) {
    const char* q = base + f.n / T * t;
    if (q < cut[t - 1]) q = cut[t - 1];
    while (q < end && !is_space(*q)) ++q;  // never cut inside a token
    cut[t] = q;
  }
  std::vector<std::vector<int>> part(T);
  std::vector<char> stopped(T, 0);
  auto work = [&](int t) {
    const char *p = cut[t], *e = cut[t + 1];
    std::vector<int>& v = part[t];
    v.reserve((size_t)(e - p) / 4 + 1);
    while (p < e) {
      while (p < e && is_space(*p)) ++p;
      if (p >= e) break;
      bool neg = false;
      if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
      if (p >= e || *p < '0' || *p > '9') { stopped[t] = 1; break; }
      long long x = 0;
      while (p < e && *p >= '0' && *p <= '9') { x = x * 10 + (*p - '0'); ++p; }
      v.push_back((int)(neg ? -x : x));
    }
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  size_t total = 0;
  for (int t = 0; t < T; ++t) {
    total += part[t].size();
    if (stopped[t]) break;
  }
  out.reserve(out.size() + total);
  for (int t = 0; t < T; ++t) {
    out.insert(out.end(), part[t].begin(), part[t].end());
    if (stopped[t]) break;
  }
}

bool read_bin(const std::string& path, std::vector<float>& out, size_t n) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  out.resize(n);
  const size_t got = fread(out.data(), sizeof(float), n, f);
  fclose(f);
  return got == n;
}
bool write_bin(const std::string& path, const std::vector<float>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const size_t put = fwrite(v.data(), sizeof(float), v.size(), f);
  fclose(f);
  return put == v.size();
}

// ---- binary dataset cache: the text format stays the interface, parsing it happens once -------------
struct CacheHeader {
  char magic[8];
  int64_t n_nodes, in_dim, n_edges;
  int64_t fsize[4], fmtime[4];
};
bool stat_files(const std::string& dir, CacheHeader& h) {
  const char* names[4] = {"features.txt", "row_ptr.txt", "col_idx.txt", "labels.txt"};
  for (int i = 0; i < 4; ++i) {
    struct stat st;
    if (stat((dir + names[i]).c_str(), &st) != 0) return false;
    h.fsize[i] = (int64_t)st.st_size;
    h.fmtime[i] = (int64_t)st.st_mtime;
  }
  return true;
}
bool cache_load(const std::string& dir, std::vector<float>& X, std::vector<int>& rp, std::vector<int>& ci,
                std::vector<int>& lab, int& N, int& I) {
  CacheHeader want{}, got{};
  if (!stat_files(dir, want)) return false;
  FILE* f = fopen((dir + ".gatx_cache.bin").c_str(), "rb");
  if (!f) return false;
  bool ok = fread(&got, sizeof got, 1, f) == 1 && memcmp(got.magic, "GATXDS1", 8) == 0 &&
            memcmp(got.fsize, want.fsize, sizeof want.fsize) == 0 && memcmp(got.fmtime, want.fmtime, sizeof want.fmtime) == 0;
  if (ok) {
    X.resize((size_t)got.n_nodes * got.in_dim);
    rp.resize((size_t)got.n_nodes + 1);
    ci.resize((size_t)got.n_edges);
    lab.resize((size_t)got.n_nodes);
    ok = fread(X.data(), sizeof(float), X.size(), f) == X.size() && fread(rp.data(), sizeof(int), rp.size(), f) == rp.size() &&
         fread(ci.data(), sizeof(int), ci.size(), f) == ci.size() && fread(lab.data(), sizeof(int), lab.size(), f) == lab.size();
    N = (int)got.n_nodes;
    I = (int)got.in_dim;
  }
  fclose(f);
  return ok;
}
void cache_store(const std::string& dir, const std::vector<float>& X, const std::vector<int>& rp,
                 const std::vector<int>& ci, const std::vector<int>& lab, int N, int I) {
  CacheHeader h{};
  memcpy(h.magic, "GATXDS1", 8);
  h.n_nodes = N; h.in_dim = I; h.n_edges = (int64_t)ci.size();
  if (!stat_files(dir, h)) return;
  const std::string tmp = dir + ".gatx_cache.bin.tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;  // read-only dataset directory: just skip the cache
  const bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(X.data(), sizeof(float), X.size(), f) == X.size() &&
                  fwrite(rp.data(), sizeof(int), rp.size(), f) == rp.size() &&
                  fwrite(ci.data(), sizeof(int), ci.size(), f) == ci.size() &&
                  fwrite(lab.data(), sizeof(int), lab.size(), f) == lab.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), (dir + ".gatx_cache.bin").c_str());
  else remove(tmp.c_str());
}

struct Args {
  int epochs = 200, L = 2, gpus = 1, gemm = GATX_GEMM_TF32_TC, every = 1;
  bool clip = false, use_cache = true, load_only = false, split = false, eval_only = false, bias = false;
  bool check_replicas = false;
  std::string save_ckpt, resume_ckpt;
  std::string optimizer = "sgd", dataset = "pubmed", data_root = "./data", load_w, dump_w;
  float lr = 0.0001f, beta1 = 0.9f, beta2 = 0.999f;
  // opt-in extensions; the defaults are the reference's fixed behaviour (slopes 0.01 / 0.01, no dropout)
  float attn_slope = 0.01f, act_slope = 0.01f, dropout = 0.0f, attn_dropout = 0.0f;
  unsigned long long seed = 0;
  bool seed_given = false;
  std::vector<int> heads, outdims;
};

// Checkpoint header (version 2): enough of the model / optimizer description to refuse a file made for another run.
constexpr int kCkptMaxLayers = 16;
struct CkptHeader {
  char magic[8];
  int32_t version, num_layers, in_dim, num_classes, use_bias, optimizer;
  int32_t heads[kCkptMaxLayers], outdims[kCkptMaxLayers];
  int64_t epochs_done, n_floats;
};

int fail_ctx(gatx_ctx* c, const char* what, int rc) {
  std::cerr << "Error: " << what << " failed (" << rc << "): " << gatx_last_error(c) << std::endl;
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  bool beta1_specified = false, beta2_specified = false;
  // the reference pre-scans --num-layers so that the lists can be sized (EB:942-953)
  for (int i = 1; i < argc; i++) {
    std::string arg = argv[i];
    if (arg == "--num-layers" && i + 1 < argc) {
      a.L = std::atoi(argv[++i]);
      if (a.L <= 0) {
        std::cerr << "Error: Number of layers must be > 0\n";
        return 1;
      }
      break;
    }
  }
  bool have_heads = false, have_outdims = false;
  for (int i = 1; i < argc; i++) {
    std::string arg = argv[i];
    auto parse_list = [&](const char* flagname, std::vector<int>& dst) -> bool {
      std::stringstream ss(argv[++i]);
      std::string item;
      dst.assign(a.L, 0);
      for (int l = 0; l < a.L; ++l) {
        if (!std::getline(ss, item, ',')) {
          std::cerr << "Error: " << flagname << " must have " << a.L << " values.\n";
          return false;
        }
        dst[l] = std::atoi(item.c_str());
      }
      return true;
    };
    if (arg == "--epochs" && i + 1 < argc) a.epochs = std::atoi(argv[++i]);
    else if (arg == "--heads" && i + 1 < argc) { if (!parse_list("--heads", a.heads)) return 1; have_heads = true; }
    else if (arg == "--outdims" && i + 1 < argc) { if (!parse_list("--ooutdims", a.outdims)) return 1; have_outdims = true; }
    else if (arg == "--clip") a.clip = true;
    else if (arg == "--optimizer" && i + 1 < argc) {
      a.optimizer = argv[++i];
      if (a.optimizer != "sgd" && a.optimizer != "adam") {
        std::cerr << "Invalid optimizer choice. Use 'sgd' or 'adam'\n";
        return 1;
      }
    }
    else if (arg == "--beta1" && i + 1 < argc) { a.beta1 = std::strtof(argv[++i], nullptr); beta1_specified = true; }
    else if (arg == "--beta2" && i + 1 < argc) { a.beta2 = std::strtof(argv[++i], nullptr); beta2_specified = true; }
    else if (arg == "--lr" && i + 1 < argc) a.lr = std::strtof(argv[++i], nullptr);
    else if (arg == "--dataset" && i + 1 < argc) a.dataset = argv[++i];
    else if (arg == "--data-root" && i + 1 < argc) a.data_root = argv[++i];
    else if (arg == "--seed" && i + 1 < argc) { a.seed = std::strtoull(argv[++i], nullptr, 10); a.seed_given = true; }
    else if (arg == "--gpus" && i + 1 < argc) a.gpus = std::max(1, std::atoi(argv[++i]));
    else if (arg == "--gemm" && i + 1 < argc) { const std::string m = argv[++i]; a.gemm = m == "fp32" ? GATX_GEMM_FP32_SIMT : (m == "3xtf32" ? GATX_GEMM_3XTF32_TC : GATX_GEMM_TF32_TC); }
    else if (arg == "--load-weights" && i + 1 < argc) a.load_w = argv[++i];
    else if (arg == "--dump-weights" && i + 1 < argc) a.dump_w = argv[++i];
    else if (arg == "--quiet-epochs" && i + 1 < argc) a.every = std::max(1, std::atoi(argv[++i]));
    else if (arg == "--save-checkpoint" && i + 1 < argc) a.save_ckpt = argv[++i];
    else if (arg == "--resume" && i + 1 < argc) a.resume_ckpt = argv[++i];
    else if (arg == "--no-cache") a.use_cache = false;
    else if (arg == "--load-only") a.load_only = true;
    else if (arg == "--split") a.split = true;
    else if (arg == "--eval-only") a.eval_only = true;
    else if (arg == "--attn-slope" && i + 1 < argc) a.attn_slope = std::strtof(argv[++i], nullptr);
    else if (arg == "--act-slope" && i + 1 < argc) a.act_slope = std::strtof(argv[++i], nullptr);
    else if (arg == "--dropout" && i + 1 < argc) a.dropout = std::strtof(argv[++i], nullptr);
    else if (arg == "--attn-dropout" && i + 1 < argc) a.attn_dropout = std::strtof(argv[++i], nullptr);
    else if (arg == "--bias") a.bias = true;
    else if (arg == "--check-replicas") a.check_replicas = true;
    // anything else is ignored, like the reference
  }
  if (!have_heads || !have_outdims) {
    // the reference reads uninitialised memory here; refuse instead
    std::cerr << "Error: --heads and --outdims must be given (" << a.L << " comma-separated values each).\n";
    return 1;
  }
  if (a.optimizer == "adam") {
    if (a.beta1 <= 0.0f || a.beta1 >= 1.0f || a.beta2 <= 0.0f || a.beta2 >= 1.0f) {
      std::cerr << "Error: For Adam optimizer, beta1 and beta2 must be in (0,1).\n";
      return 1;
    }
  } else if (beta1_specified || beta2_specified) {
    std::cerr << "Warning: beta1/beta2 specified but ignored for SGD optimizer.\n";
  }

  std::cout << "Configuration:\n"
            << "  Number of layers: " << a.L << "\n"
            << "  Epochs: " << a.epochs << "\n"
            << "  Attention heads: [";
  for (int l = 0; l < a.L; ++l) std::cout << a.heads[l] << (l < a.L - 1 ? ", " : "");
  std::cout << "]\n  Output dimensions: [";
  for (int l = 0; l < a.L; ++l) std::cout << a.outdims[l] << (l < a.L - 1 ? ", " : "");
  std::cout << "]\n"
            << "  Gradient clipping: " << (a.clip ? "true" : "false") << "\n"
            << "  Optimizer: " << a.optimizer << "\n"
            << "  Learning rate: " << a.lr << "\n\n";

  const char* env_root = std::getenv("DATA_ROOT");
  if (env_root && a.data_root == "./data") a.data_root = env_root;  // EB:1064-1067
  if (!a.data_root.empty() && a.data_root.back() != '/' && a.data_root.back() != '\\') a.data_root += '/';
  const std::string path = a.data_root + a.dataset + "/";
  std::cout << "Using dataset: " << a.dataset << std::endl;
  std::cout << "Dataset path: " << path << std::endl;

  std::vector<float> X;
  std::vector<int> row_ptr, col_idx, labels;
  int N = 0, I = 0;
  const auto t_load0 = std::chrono::high_resolution_clock::now();
  const bool from_cache = a.use_cache && cache_load(path, X, row_ptr, col_idx, labels, N, I);
  if (!from_cache) {
    load_features(path + "features.txt", X, N, I);
    load_int_array(path + "row_ptr.txt", row_ptr);
  }
  if ((int)row_ptr.size() != N + 1) {
    std::cerr << "Invalid row_ptr length\n";
    return 1;
  }
  if (!from_cache) {
    load_int_array(path + "col_idx.txt", col_idx);
    load_int_array(path + "labels.txt", labels);
  }
  if ((int)labels.size() != N) {
    std::cerr << "Invalid labels length\n";
    return 1;
  }
  if (N == 0 || (long long)row_ptr[N] != (long long)col_idx.size()) {
    std::cerr << "Invalid col_idx length\n";
    return 1;
  }
  int max_degree = 0;
  for (int i = 0; i < N; ++i) max_degree = std::max(max_degree, row_ptr[i + 1] - row_ptr[i]);
  for (int i = 0; i < N; ++i)
    if (labels[i] < 0) {  // the reference would index its class arrays out of bounds (EB:524, EB:572)
      std::cerr << "Invalid label on line " << i + 1 << "\n";
      return 1;
    }
  const int C = *std::max_element(labels.begin(), labels.end()) + 1;
  std::cout << "Max degree = " << max_degree << std::endl;
  std::cout << "Number of classes = " << C << std::endl;
  std::cout << "Graph loaded: " << N << " nodes, " << col_idx.size() << " edges, "
            << "input_feature_vector_dim = " << I << std::endl;
  if (a.use_cache && !from_cache) cache_store(path, X, row_ptr, col_idx, labels, N, I);
  {
    const std::chrono::duration<double, std::milli> dt = std::chrono::high_resolution_clock::now() - t_load0;
    std::cerr << "[loader] " << (from_cache ? "binary cache" : "text files") << ": " << dt.count() << " ms\n";
  }
  if (a.load_only) return 0;
  std::vector<unsigned char> mask[3];  // train / validation / test
  if (a.split) {
    std::vector<int> part;
    load_int_array(path + "split.txt", part);
    if ((int)part.size() != N) {
      std::cerr << "Invalid split length\n";
      return 1;
    }
    for (int k = 0; k < 3; ++k) mask[k].assign((size_t)N, 0);
    for (int i = 0; i < N; ++i) {
      if (part[i] < 0 || part[i] > 2) {
        std::cerr << "Invalid split value on line " << i + 1 << " (0 = train, 1 = validation, 2 = test)\n";
        return 1;
      }
      mask[part[i]][i] = 1;
    }
  }

  const int world = a.gpus;
  std::vector<gatx_ctx*> ctx(world, nullptr);
  unsigned char nccl_id[128];
  if (world > 1 && gatx_comm_unique_id(nccl_id) != GATX_OK) {
    std::cerr << "Error: NCCL is not available for --gpus " << world << "\n";
    return 1;
  }
  {
    int ndev = gatx_device_count();
    if (ndev < world) {  // checked here: a rank that cannot be created would leave the others waiting in NCCL
      std::cerr << "Error: --gpus " << world << " but " << (ndev < 0 ? 0 : ndev) << " usable CUDA device(s)\n";
      return 1;
    }
  }
  // the reference seeds with time(NULL) (EB:1305); --seed makes runs reproducible.  ONE value for every rank: the
  // parameters are replicated, never broadcast.
  const unsigned long long init_seed = a.seed_given ? a.seed : (unsigned long long)time(nullptr);
  std::atomic<int> failed{0};
  auto for_ranks = [&](auto&& fn) {
    if (world == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r) th.emplace_back(fn, r);
    for (auto& t : th) t.join();
  };
  // phase 0: contexts, one after the other (nothing collective yet)
  for (int r = 0; r < world; ++r) {
    gatx_config cfg{};
    cfg.num_layers = a.L;
    cfg.heads = a.heads.data();
    cfg.outdims = a.outdims.data();
    cfg.optimizer = a.optimizer == "adam" ? GATX_OPT_ADAM : GATX_OPT_SGD;
    cfg.lr = a.lr; cfg.beta1 = a.beta1; cfg.beta2 = a.beta2;
    cfg.clip = a.clip ? 1 : 0;
    cfg.device = r; cfg.gemm_mode = a.gemm; cfg.keep_debug = 0; cfg.rank = r; cfg.world = world;
    int rc = gatx_create(&ctx[r], &cfg);
    if (rc) {
      std::cerr << "Error: gatx_create failed (" << rc << ") on device " << r << " -- no usable sm_100 GPU?\n";
      for (auto c : ctx) gatx_destroy(c);
      return 1;
    }
  }
  // phase 1: everything a rank does on its own (graph preparation, labels, masks, options)
  for_ranks([&](int r) {
    int rc;
    if ((rc = gatx_set_graph_csr(ctx[r], N, (int64_t)col_idx.size(), row_ptr.data(), col_idx.data()))) { failed = fail_ctx(ctx[r], "gatx_set_graph_csr", rc); return; }
    if ((rc = gatx_set_labels(ctx[r], labels.data(), C))) { failed = fail_ctx(ctx[r], "gatx_set_labels", rc); return; }
    if (a.split && (rc = gatx_set_train_mask(ctx[r], mask[0].data()))) { failed = fail_ctx(ctx[r], "gatx_set_train_mask", rc); return; }
    if (a.bias && (rc = gatx_set_bias(ctx[r], 1))) { failed = fail_ctx(ctx[r], "gatx_set_bias", rc); return; }
    if ((a.attn_slope != 0.01f || a.act_slope != 0.01f) && (rc = gatx_set_slopes(ctx[r], a.attn_slope, a.act_slope))) { failed = fail_ctx(ctx[r], "gatx_set_slopes", rc); return; }
  });
  if (failed) {  // no rank has entered a collective yet: a clean exit
    for (auto c : ctx) gatx_destroy(c);
    return 1;
  }
  // phase 2: the collective part (communicator, feature all-gather) and the parameters.  A rank that fails here would
  // leave its peers blocked inside NCCL, so the process ends at once instead of joining them.
  auto die = [&](gatx_ctx* c, const char* what, int rc) {
    fail_ctx(c, what, rc);
    if (world > 1) std::_Exit(1);
    failed = 1;
  };
  for_ranks([&](int r) {
    int rc;
    if (world > 1 && (rc = gatx_comm_init(ctx[r], nccl_id))) { die(ctx[r], "gatx_comm_init", rc); return; }
    if ((rc = gatx_set_features(ctx[r], X.data(), I))) { die(ctx[r], "gatx_set_features", rc); return; }
    if ((rc = gatx_init_params(ctx[r], init_seed))) { die(ctx[r], "gatx_init_params", rc); return; }
    if (!a.load_w.empty()) {
      std::vector<float> W, av, Wo;
      size_t wo = 0, ao = 0, tw = 0, ta = 0;
      int in = I;
      for (int l = 0; l < a.L; ++l) { tw += (size_t)a.heads[l] * a.outdims[l] * 2 * in; ta += (size_t)a.heads[l] * a.outdims[l]; in = a.heads[l] * a.outdims[l]; }
      if (!read_bin(a.load_w + "/W.bin", W, tw) || !read_bin(a.load_w + "/a.bin", av, ta) ||
          !read_bin(a.load_w + "/Wo.bin", Wo, (size_t)C * a.outdims[a.L - 1])) {
        std::cerr << "Error: cannot read W.bin / a.bin / Wo.bin from " << a.load_w << "\n";
        if (world > 1) std::_Exit(1);
        failed = 1;
        return;
      }
      in = I;
      for (int l = 0; l < a.L; ++l) {
        if ((rc = gatx_set_params(ctx[r], l, W.data() + wo, av.data() + ao))) { die(ctx[r], "gatx_set_params", rc); return; }
        wo += (size_t)a.heads[l] * a.outdims[l] * 2 * in;
        ao += (size_t)a.heads[l] * a.outdims[l];
        in = a.heads[l] * a.outdims[l];
      }
      if ((rc = gatx_set_wo(ctx[r], Wo.data()))) { die(ctx[r], "gatx_set_wo", rc); return; }
      if (a.bias) {  // b.bin: the per-layer biases [b_0 | .. | b_{L-1}]; absent in dumps made without --bias (they stay zero)
        std::vector<float> bv;
        if (read_bin(a.load_w + "/b.bin", bv, ta)) {
          size_t bo = 0;
          for (int l = 0; l < a.L; ++l) {
            if ((rc = gatx_set_bias_values(ctx[r], l, bv.data() + bo))) { die(ctx[r], "gatx_set_bias_values", rc); return; }
            bo += (size_t)a.heads[l] * a.outdims[l];
          }
        }
      }
    }
  });
  if (failed) return 1;
  if (world > 1) {
    // halo exchange over NVLink peer memory: the contexts live in one process, so the blobs carry raw pointers.  The
    // decision is GLOBAL: if any rank cannot export or import, every rank falls back to the NCCL collectives (ranks
    // issuing different exchange sequences would deadlock).
    std::vector<unsigned char> blobs((size_t)world * GATX_PEER_INFO_BYTES);
    bool ok = true;
    int bad = 0;
    for (int r = 0; r < world && ok; ++r)
      if (gatx_peer_export(ctx[r], blobs.data() + (size_t)r * GATX_PEER_INFO_BYTES, GATX_PEER_INFO_BYTES) != GATX_OK) { ok = false; bad = r; }
    for (int r = 0; r < world && ok; ++r)
      if (gatx_peer_import(ctx[r], blobs.data(), blobs.size()) != GATX_OK) { ok = false; bad = r; }
    if (!ok) {
      std::cerr << "Note: peer-memory halo exchange unavailable (" << gatx_last_error(ctx[bad]) << "); using NCCL collectives\n";
      for (int r = 0; r < world; ++r) gatx_peer_disable(ctx[r]);
    }
  }

  auto dump_weights = [&](const std::string& dir) {
    std::vector<float> W, av, Wo;
    for (int l = 0; l < a.L; ++l) {
      const int64_t nw = gatx_tensor_size(ctx[0], GATX_T_W, l), na = gatx_tensor_size(ctx[0], GATX_T_A, l);
      std::vector<float> w((size_t)nw), v((size_t)na);
      gatx_get_tensor(ctx[0], GATX_T_W, l, w.data(), w.size() * 4);
      gatx_get_tensor(ctx[0], GATX_T_A, l, v.data(), v.size() * 4);
      W.insert(W.end(), w.begin(), w.end());
      av.insert(av.end(), v.begin(), v.end());
    }
    Wo.resize((size_t)gatx_tensor_size(ctx[0], GATX_T_WO, 0));
    gatx_get_tensor(ctx[0], GATX_T_WO, 0, Wo.data(), Wo.size() * 4);
    if (a.bias) {
      std::vector<float> bv;
      for (int l = 0; l < a.L; ++l) {
        std::vector<float> b((size_t)gatx_tensor_size(ctx[0], GATX_T_B, l));
        gatx_get_tensor(ctx[0], GATX_T_B, l, b.data(), b.size() * 4);
        bv.insert(bv.end(), b.begin(), b.end());
      }
      if (!write_bin(dir + "/b.bin", bv)) return false;
    }
    return write_bin(dir + "/W.bin", W) && write_bin(dir + "/a.bin", av) && write_bin(dir + "/Wo.bin", Wo);
  };

  // checkpoint file: CkptHeader (magic, version, the model / optimizer description), then [params | Adam m | Adam v].
  // --resume refuses a file written for another architecture even when the parameter COUNT happens to match
  // (4 heads x 64 and 8 heads x 32 have the same number of floats).
  CkptHeader want{};
  memcpy(want.magic, "GATXCK2", 8);
  want.version = 2;
  want.num_layers = a.L;
  want.in_dim = I;
  want.num_classes = C;
  want.use_bias = a.bias ? 1 : 0;
  want.optimizer = a.optimizer == "adam" ? GATX_OPT_ADAM : GATX_OPT_SGD;
  for (int l = 0; l < a.L && l < kCkptMaxLayers; ++l) { want.heads[l] = a.heads[l]; want.outdims[l] = a.outdims[l]; }
  want.n_floats = gatx_state_size(ctx[0]);
  int first_epoch = 1;
  if (!a.resume_ckpt.empty()) {
    FILE* f = fopen(a.resume_ckpt.c_str(), "rb");
    CkptHeader got{};
    std::vector<float> stt;
    bool ok = f && fread(&got, sizeof got, 1, f) == 1;
    const char* why = "missing or not a gatx checkpoint";
    if (ok && (memcmp(got.magic, want.magic, 8) != 0 || got.version != want.version)) ok = false;
    if (ok) {
      why = "written for another model";
      ok = a.L <= kCkptMaxLayers && got.num_layers == want.num_layers && got.in_dim == want.in_dim &&
           got.num_classes == want.num_classes && got.use_bias == want.use_bias && got.n_floats == want.n_floats &&
           memcmp(got.heads, want.heads, sizeof want.heads) == 0 && memcmp(got.outdims, want.outdims, sizeof want.outdims) == 0;
    }
    if (ok && got.optimizer != want.optimizer) { ok = false; why = "written by another optimizer"; }
    if (ok) {
      why = "truncated";
      stt.resize((size_t)got.n_floats);
      ok = fread(stt.data(), sizeof(float), stt.size(), f) == stt.size();
    }
    if (f) fclose(f);
    if (!ok) {
      std::cerr << "Error: cannot resume from " << a.resume_ckpt << " (" << why << ")\n";
      return 1;
    }
    for (int r = 0; r < world; ++r)
      if (int rc = gatx_set_state(ctx[r], stt.data(), stt.size() * sizeof(float))) return fail_ctx(ctx[r], "gatx_set_state", rc);
    first_epoch = (int)got.epochs_done + 1;
  }
  if (a.dropout != 0.0f) {
    // every rank draws the same mask (a function of seed, layer, step, global row, column); a resumed run continues
    // with a fresh stream
    const unsigned long long dseed = init_seed + 7919ull * (unsigned long long)first_epoch;
    for (int r = 0; r < world; ++r)
      if (int rc = gatx_set_dropout(ctx[r], a.dropout, dseed)) return fail_ctx(ctx[r], "gatx_set_dropout", rc);
    std::cout << "Dropout: " << a.dropout << " on every layer's input (training forwards only)\n";
  }
  if (a.attn_dropout != 0.0f) {
    const unsigned long long aseed = init_seed + 104729ull * (unsigned long long)first_epoch;
    for (int r = 0; r < world; ++r)
      if (int rc = gatx_set_attn_dropout(ctx[r], a.attn_dropout, aseed)) return fail_ctx(ctx[r], "gatx_set_attn_dropout", rc);
    std::cout << "Attention dropout: " << a.attn_dropout << " on the attention coefficients (training forwards only)\n";
  }
  // evaluation forward on every rank; returns rank 0's (already all-reduced) scalars
  auto evaluate = [&](const unsigned char* m, float* lo, float* ac) -> int {
    std::vector<float> l2(world, 0.f), a2(world, 0.f);
    std::vector<int> rcs(world, 0);
    auto run = [&](int r) { rcs[r] = gatx_evaluate(ctx[r], m, &l2[r], &a2[r]); };
    if (world == 1) run(0);
    else {
      std::vector<std::thread> th;
      for (int r = 0; r < world; ++r) th.emplace_back(run, r);
      for (auto& t : th) t.join();
    }
    for (int r = 0; r < world; ++r)
      if (rcs[r]) return fail_ctx(ctx[r], "gatx_evaluate", rcs[r]);
    *lo = l2[0];
    *ac = a2[0];
    return 0;
  };
  if (a.eval_only) {
    float lo = 0.f, ac = 0.f;
    if (!a.split) {
      if (evaluate(nullptr, &lo, &ac)) return 1;
      printf("\nAvg Loss: %f, Accuracy: %.2f%%\n", lo, 100.0f * ac);
    } else {
      const char* names[3] = {"Train", "Val", "Test"};
      for (int k = 0; k < 3; ++k) {
        if (evaluate(mask[k].data(), &lo, &ac)) return 1;
        printf("\n%s Loss: %f, %s Accuracy: %.2f%%\n", names[k], lo, names[k], 100.0f * ac);
      }
    }
    for (auto c : ctx) gatx_destroy(c);
    return 0;
  }
  const int last_epoch = first_epoch + a.epochs - 1;
  for (int epoch = first_epoch; epoch <= last_epoch; ++epoch) {
    auto start = std::chrono::high_resolution_clock::now();
    const bool show = (epoch % a.every) == 0 || epoch == first_epoch || epoch == last_epoch;
    if (show) printf("\nEpoch %d\n", epoch);
    std::vector<float> loss(world, 0.f), acc(world, 0.f);
    std::vector<int> rcs(world, 0);
    auto run = [&](int r) { rcs[r] = gatx_train_epoch(ctx[r], epoch, &loss[r], &acc[r]); };
    if (world == 1) run(0);
    else {
      std::vector<std::thread> th;
      for (int r = 0; r < world; ++r) th.emplace_back(run, r);
      for (auto& t : th) t.join();
    }
    for (int r = 0; r < world; ++r)
      if (rcs[r]) return fail_ctx(ctx[r], "gatx_train_epoch", rcs[r]);
    if (show) printf("\nAvg Loss: %f, Accuracy: %.2f%%\n", loss[0], 100.0f * acc[0]);
    if (show && a.split) {
      float lo = 0.f, ac = 0.f;
      if (evaluate(mask[1].data(), &lo, &ac)) return 1;
      printf("\nVal Loss: %f, Val Accuracy: %.2f%%\n", lo, 100.0f * ac);
    }
    auto stop = std::chrono::high_resolution_clock::now();
    std::chrono::duration<double, std::milli> elapsed = stop - start;
    if (show) std::cout << " total time: " << elapsed.count() << " ms" << std::endl;
  }
  if (a.split) {
    float lo = 0.f, ac = 0.f;
    if (evaluate(mask[2].data(), &lo, &ac)) return 1;
    printf("\nTest Loss: %f, Test Accuracy: %.2f%%\n", lo, 100.0f * ac);
  }
  if (!a.save_ckpt.empty()) {
    const int64_t n = want.n_floats;
    std::vector<float> stt((size_t)(n > 0 ? n : 0));
    CkptHeader hdr = want;
    hdr.epochs_done = last_epoch;
    FILE* f = n > 0 && a.L <= kCkptMaxLayers && gatx_get_state(ctx[0], stt.data(), stt.size() * sizeof(float)) == GATX_OK
                  ? fopen(a.save_ckpt.c_str(), "wb") : nullptr;
    if (!f || fwrite(&hdr, sizeof hdr, 1, f) != 1 || fwrite(stt.data(), sizeof(float), stt.size(), f) != stt.size())
      std::cerr << "Warning: could not write checkpoint " << a.save_ckpt << "\n";
    if (f) fclose(f);
  }
  if (a.check_replicas && world > 1) {
    // every rank applies the same all-reduced gradients to its own copy of the parameters: the copies must be identical
    const int64_t n = gatx_state_size(ctx[0]);
    std::vector<float> s0((size_t)n), sr((size_t)n);
    bool same = n > 0 && gatx_get_state(ctx[0], s0.data(), s0.size() * sizeof(float)) == GATX_OK;
    for (int r = 1; r < world && same; ++r)
      same = gatx_get_state(ctx[r], sr.data(), sr.size() * sizeof(float)) == GATX_OK &&
             memcmp(s0.data(), sr.data(), s0.size() * sizeof(float)) == 0;
    std::cout << (same ? "Replicas identical on " : "REPLICAS DIFFER on ") << world << " ranks" << std::endl;
    if (!same) return 1;
  }
  if (!a.dump_w.empty() && !dump_weights(a.dump_w)) std::cerr << "Warning: could not write weights to " << a.dump_w << "\n";
  for (auto c : ctx) gatx_destroy(c);
  return 0;
}

"""ctypes binding of libgatx.so (include/gatx.h) -- the call surface tests and bench.py use.

There is deliberately no fallback: if the CUDA library is missing, cannot be built, or no
sm_100 device is present, every entry point raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# GATX_LIB selects an A/B variant built by `GATX_VARIANT=... build.py` (tools/ab_packed_fp32.sh); default: the product
LIB_PATH = os.environ.get("GATX_LIB") or os.path.join(HERE, "libgatx.so")

(T_W, T_A, T_WO, T_GW, T_GA, T_GWO, T_PL, T_PR, T_SCORE, T_ALPHA, T_HPRE, T_HOUT, T_Y, T_GH, T_Z, T_PRED,
 T_COO_SRC, T_COO_DST, T_IN_DEGREE, T_CSC_PTR, T_CSC_DST, T_CSC_EID, T_GPL, T_GPR, T_GALPHA, T_GE, T_B, T_GB) = range(28)
_INT_TENSORS = {T_PRED, T_COO_SRC, T_COO_DST, T_IN_DEGREE, T_CSC_PTR, T_CSC_DST, T_CSC_EID}
GEMM_TF32_TC, GEMM_FP32_SIMT, GEMM_3XTF32_TC = 0, 1, 2
PEER_INFO_BYTES = 2048
PHASES = ("gemm_fwd", "edge_fwd", "head", "edge_bwd", "gemm_bwd", "optimizer", "comm", "epoch")

EXPORTS = [
    "gatx_create", "gatx_destroy", "gatx_last_error", "gatx_version", "gatx_set_graph_csr",
    "gatx_set_features", "gatx_set_labels", "gatx_graph_info", "gatx_partition_rows", "gatx_row_blocks", "gatx_init_params",
    "gatx_set_params", "gatx_set_wo", "gatx_forward", "gatx_loss_acc", "gatx_backward", "gatx_step",
    "gatx_train_epoch", "gatx_sync", "gatx_tensor_size", "gatx_get_tensor", "gatx_enable_timing",
    "gatx_get_timing", "gatx_get_edge_kernel_ms", "gatx_timer_start", "gatx_timer_stop", "gatx_launch_count", "gatx_edge_bytes", "gatx_state_size", "gatx_get_state", "gatx_set_state", "gatx_set_train_mask", "gatx_evaluate", "gatx_op_gemm", "gatx_op_edge_fwd", "gatx_op_edge_bwd", "gatx_op_softmax_ce", "gatx_op_optimizer",
    "gatx_comm_unique_id", "gatx_comm_init", "gatx_peer_export", "gatx_peer_import", "gatx_halo_rows", "gatx_halo_active",
    "gatx_device_count", "gatx_peer_disable", "gatx_halo_stats", "gatx_set_cuda_graph", "gatx_cuda_graph_active", "gatx_set_slopes", "gatx_set_dropout", "gatx_set_attn_dropout", "gatx_set_bias", "gatx_set_bias_values",
]


class GatxConfig(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("heads", C.POINTER(C.c_int32)), ("outdims", C.POINTER(C.c_int32)),
                ("optimizer", C.c_int32), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("clip", C.c_int32), ("device", C.c_int32), ("gemm_mode", C.c_int32), ("keep_debug", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32)]


class GatxError(RuntimeError):
    pass


_lib = None


def load(build_if_missing=False):
    """Loads libgatx.so; raises if it is not there (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise GatxError("libgatx.so is not built: run `python graph-attention-network-gatv2-_b200/build.py` "
                            "(there is no CPU fallback)")
        import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.gatx_last_error.restype = C.c_char_p
    lib.gatx_version.restype = C.c_char_p
    lib.gatx_tensor_size.restype = C.c_int64
    lib.gatx_launch_count.restype = C.c_int64
    lib.gatx_destroy.restype = None
    lib.gatx_set_graph_csr.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.gatx_set_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.gatx_set_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.gatx_init_params.argtypes = [C.c_void_p, C.c_uint64]
    lib.gatx_set_params.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gatx_set_wo.argtypes = [C.c_void_p, C.c_void_p]
    lib.gatx_get_tensor.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
    lib.gatx_tensor_size.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.gatx_train_epoch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gatx_loss_acc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gatx_step.argtypes = [C.c_void_p, C.c_int32]
    lib.gatx_timer_stop.argtypes = [C.c_void_p, C.c_void_p]
    for fn in ("gatx_forward", "gatx_backward", "gatx_sync", "gatx_destroy", "gatx_last_error", "gatx_launch_count",
               "gatx_timer_start"):
        getattr(lib, fn).argtypes = [C.c_void_p]
    lib.gatx_enable_timing.argtypes = [C.c_void_p, C.c_int32]
    lib.gatx_get_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.gatx_edge_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gatx_graph_info.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.gatx_partition_rows.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.gatx_comm_unique_id.argtypes = [C.c_void_p]
    lib.gatx_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def partition_rows(row_ptr, world):
    row_ptr = np.ascontiguousarray(row_ptr, np.int32)
    b = np.empty(world + 1, np.int32)
    rc = load().gatx_partition_rows(len(row_ptr) - 1, row_ptr.ctypes.data, world, b.ctypes.data)
    if rc:
        raise GatxError("gatx_partition_rows failed: %d" % rc)
    return b


def row_blocks(row_ptr, world, num_blocks):
    """[world][num_blocks + 1] global row bounds of every rank's exchange blocks."""
    row_ptr = np.ascontiguousarray(row_ptr, np.int32)
    out = np.empty((world, num_blocks + 1), np.int32)
    lib = load()
    lib.gatx_row_blocks.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    rc = lib.gatx_row_blocks(len(row_ptr) - 1, row_ptr.ctypes.data, world, num_blocks, out.ctypes.data)
    if rc:
        raise GatxError("gatx_row_blocks failed: %d" % rc)
    return out


def op_gemm(A, B, form=0, mode=GEMM_TF32_TC):
    """form 0: A[M][K] @ B[N][K].T ; form 1: A[K][M].T @ B[K][N] -- through the engine's GEMM kernels."""
    A, B = np.ascontiguousarray(A, np.float32), np.ascontiguousarray(B, np.float32)
    if form == 0:
        (M, K), N = A.shape, B.shape[0]
    else:
        (K, M), N = A.shape, B.shape[1]
    Cm = np.zeros((M, N), np.float32)
    lib = load()
    lib.gatx_op_gemm.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_int64, C.c_int32, C.c_int32, C.c_int64]
    rc = lib.gatx_op_gemm(mode, form, A.ctypes.data, A.shape[1], B.ctypes.data, B.shape[1], Cm.ctypes.data, N, M, N, K)
    if rc:
        raise GatxError("gatx_op_gemm failed: %d" % rc)
    return Cm


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def op_edge_fwd(row_ptr, col_idx, H, D, Pl, Pr, a):
    """Fused edge forward of one layer on host arrays -> dict(score [E][H], alpha [E][H], hpre, Hout [N][H*D])."""
    rp, ci = np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col_idx, np.int32)
    N, E, F = len(rp) - 1, len(ci), H * D
    Pl, Pr, a = _f32(Pl), _f32(Pr), _f32(a)
    out = dict(score=np.zeros((E, H), np.float32), alpha=np.zeros((E, H), np.float32), hpre=np.zeros((N, F), np.float32),
               Hout=np.zeros((N, F), np.float32))
    lib = load()
    lib.gatx_op_edge_fwd.argtypes = [C.c_int32, C.c_int64] + [C.c_void_p] * 2 + [C.c_int32] * 2 + [C.c_void_p] * 7
    rc = lib.gatx_op_edge_fwd(N, E, rp.ctypes.data, ci.ctypes.data, H, D, Pl.ctypes.data, Pr.ctypes.data, a.ctypes.data,
                              out["score"].ctypes.data, out["alpha"].ctypes.data, out["hpre"].ctypes.data,
                              out["Hout"].ctypes.data)
    if rc:
        raise GatxError("gatx_op_edge_fwd failed: %d" % rc)
    return out


def op_edge_bwd(row_ptr, col_idx, H, D, Pl, Pr, a, gHout):
    """Fused edge backward of one layer on host arrays -> dict(g_pre, gPl, gPr [N][H*D], ga [H*D], ge [E][H])."""
    rp, ci = np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col_idx, np.int32)
    N, E, F = len(rp) - 1, len(ci), H * D
    Pl, Pr, a, gHout = _f32(Pl), _f32(Pr), _f32(a), _f32(gHout)
    out = dict(g_pre=np.zeros((N, F), np.float32), gPl=np.zeros((N, F), np.float32), gPr=np.zeros((N, F), np.float32),
               ga=np.zeros(F, np.float32), ge=np.zeros((E, H), np.float32))
    lib = load()
    lib.gatx_op_edge_bwd.argtypes = [C.c_int32, C.c_int64] + [C.c_void_p] * 2 + [C.c_int32] * 2 + [C.c_void_p] * 9
    rc = lib.gatx_op_edge_bwd(N, E, rp.ctypes.data, ci.ctypes.data, H, D, Pl.ctypes.data, Pr.ctypes.data, a.ctypes.data,
                              gHout.ctypes.data, out["g_pre"].ctypes.data, out["gPl"].ctypes.data, out["gPr"].ctypes.data,
                              out["ga"].ctypes.data, out["ge"].ctypes.data)
    if rc:
        raise GatxError("gatx_op_edge_bwd failed: %d" % rc)
    return out


def op_softmax_ce(z, labels, mask=None):
    """Softmax + CE + argmax + dz on logits [N][C] -> dict(y, dz, pred, loss_sum, correct)."""
    z, labels = _f32(z), np.ascontiguousarray(labels, np.int32)
    N, Cc = z.shape
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    y, dz, pred = np.zeros_like(z), np.zeros_like(z), np.zeros(N, np.int32)
    ls, cc = C.c_double(), C.c_int64()
    lib = load()
    lib.gatx_op_softmax_ce.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 8
    rc = lib.gatx_op_softmax_ce(N, Cc, z.ctypes.data, labels.ctypes.data, None if m is None else m.ctypes.data,
                                y.ctypes.data, dz.ctypes.data, pred.ctypes.data, C.addressof(ls), C.addressof(cc))
    if rc:
        raise GatxError("gatx_op_softmax_ce failed: %d" % rc)
    return dict(y=y, dz=dz, pred=pred, loss_sum=ls.value, correct=cc.value)


def op_optimizer(params, grads, group_ends, optimizer="adam", clip=False, lr=1e-3, beta1=0.9, beta2=0.999, t=1, m=None, v=None):
    """One clip + Adam / SGD step on flat host vectors (returned updated: params, grads (zeroed), m, v)."""
    p, g = _f32(params).copy(), _f32(grads).copy()
    n = len(p)
    adam = optimizer == "adam"
    mm = (np.zeros(n, np.float32) if m is None else _f32(m).copy()) if adam else None
    vv = (np.zeros(n, np.float32) if v is None else _f32(v).copy()) if adam else None
    ge = (C.c_int64 * 3)(*[int(x) for x in group_ends])
    lib = load()
    lib.gatx_op_optimizer.argtypes = [C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32] + [C.c_void_p] * 4
    rc = lib.gatx_op_optimizer(n, ge, 1 if adam else 0, int(clip), lr, beta1, beta2, t, p.ctypes.data, g.ctypes.data,
                               None if mm is None else mm.ctypes.data, None if vv is None else vv.ctypes.data)
    if rc:
        raise GatxError("gatx_op_optimizer failed: %d" % rc)
    return p, g, mm, vv


def comm_unique_id():
    buf = (C.c_char * 128)()
    rc = load().gatx_comm_unique_id(buf)
    if rc:
        raise GatxError("gatx_comm_unique_id failed: %d" % rc)
    return bytes(buf)


class Engine:
    """One context (one GPU / rank).  Mirrors the reference's epoch: forward, loss_acc, backward, step."""

    def __init__(self, heads, outdims, optimizer="sgd", lr=1e-4, beta1=0.9, beta2=0.999, clip=False, device=0,
                 gemm_mode=GEMM_TF32_TC, keep_debug=False, rank=0, world=1):
        self.lib = load()
        self.heads, self.outdims = list(heads), list(outdims)
        self.L = len(self.heads)
        self._h = (C.c_int32 * self.L)(*self.heads)
        self._d = (C.c_int32 * self.L)(*self.outdims)
        cfg = GatxConfig(self.L, self._h, self._d, 1 if optimizer == "adam" else 0, lr, beta1, beta2, int(clip),
                         device, gemm_mode, int(keep_debug), rank, world)
        self.ctx = C.c_void_p()
        rc = self.lib.gatx_create(C.byref(self.ctx), C.byref(cfg))
        if rc:
            self.ctx = None
            raise GatxError("gatx_create failed with %d (no usable sm_100 CUDA device?)" % rc)
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.gatx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise GatxError("%s failed (%d): %s" % (what, rc, self.lib.gatx_last_error(self.ctx).decode()))

    # ---- data
    def set_graph(self, row_ptr, col_idx):
        self._row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        self._col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.N, self.E = len(self._row_ptr) - 1, len(self._col_idx)
        self._ck(self.lib.gatx_set_graph_csr(self.ctx, self.N, self.E, self._row_ptr.ctypes.data,
                                             self._col_idx.ctypes.data), "gatx_set_graph_csr")

    def set_features(self, X):
        X = np.ascontiguousarray(X, np.float32)
        self.I0 = X.shape[1]
        self._ck(self.lib.gatx_set_features(self.ctx, X.ctypes.data, X.shape[1]), "gatx_set_features")

    def set_features_ptr(self, ptr, in_dim):
        self._ck(self.lib.gatx_set_features(self.ctx, ptr, in_dim), "gatx_set_features")

    def set_labels(self, labels, num_classes=0):
        labels = np.ascontiguousarray(labels, np.int32)
        self._ck(self.lib.gatx_set_labels(self.ctx, labels.ctypes.data, num_classes), "gatx_set_labels")

    def set_labels_ptr(self, ptr, num_classes):
        self._ck(self.lib.gatx_set_labels(self.ctx, ptr, num_classes), "gatx_set_labels")

    def graph_info(self):
        v = [C.c_int32() for _ in range(4)]
        self._ck(self.lib.gatx_graph_info(self.ctx, *[C.byref(x) for x in v]), "gatx_graph_info")
        return dict(max_degree=v[0].value, num_classes=v[1].value, row_begin=v[2].value, row_end=v[3].value)

    # ---- parameters
    def init_params(self, seed=0):
        self._ck(self.lib.gatx_init_params(self.ctx, seed), "gatx_init_params")

    def set_params(self, layer, W, a):
        W, a = np.ascontiguousarray(W, np.float32), np.ascontiguousarray(a, np.float32)
        self._ck(self.lib.gatx_set_params(self.ctx, layer, W.ctypes.data, a.ctypes.data), "gatx_set_params")

    def set_wo(self, Wo):
        Wo = np.ascontiguousarray(Wo, np.float32)
        self._ck(self.lib.gatx_set_wo(self.ctx, Wo.ctypes.data), "gatx_set_wo")

    # ---- epoch
    def forward(self):
        self._ck(self.lib.gatx_forward(self.ctx), "gatx_forward")

    def loss_acc(self):
        a, b = C.c_float(), C.c_float()
        self._ck(self.lib.gatx_loss_acc(self.ctx, C.byref(a), C.byref(b)), "gatx_loss_acc")
        return a.value, b.value

    def backward(self):
        self._ck(self.lib.gatx_backward(self.ctx), "gatx_backward")

    def step(self, t):
        self._ck(self.lib.gatx_step(self.ctx, t), "gatx_step")

    def train_epoch(self, t, want_loss=True):
        if not want_loss:
            self._ck(self.lib.gatx_train_epoch(self.ctx, t, None, None), "gatx_train_epoch")
            return None
        a, b = C.c_float(), C.c_float()
        self._ck(self.lib.gatx_train_epoch(self.ctx, t, C.byref(a), C.byref(b)), "gatx_train_epoch")
        return a.value, b.value

    def sync(self):
        self._ck(self.lib.gatx_sync(self.ctx), "gatx_sync")

    # ---- introspection
    def tensor(self, which, layer=0):
        n = self.lib.gatx_tensor_size(self.ctx, which, layer)
        if n < 0:
            raise GatxError("tensor %d/%d not available" % (which, layer))
        out = np.empty(n, np.int32 if which in _INT_TENSORS else np.float32)
        self._ck(self.lib.gatx_get_tensor(self.ctx, which, layer, out.ctypes.data, out.nbytes), "gatx_get_tensor")
        return out

    def enable_timing(self, on=True):
        self._ck(self.lib.gatx_enable_timing(self.ctx, int(on)), "gatx_enable_timing")

    def timing(self):
        buf = (C.c_float * 8)()
        self._ck(self.lib.gatx_get_timing(self.ctx, buf, 8), "gatx_get_timing")
        return dict(zip(PHASES, list(buf)))

    def edge_kernel_ms(self, layer):
        buf = (C.c_float * 3)()
        self.lib.gatx_get_edge_kernel_ms.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        self._ck(self.lib.gatx_get_edge_kernel_ms(self.ctx, layer, buf), "gatx_get_edge_kernel_ms")
        return dict(fwd=buf[0], bwd_dst=buf[1], bwd_src=buf[2])

    def timer_start(self):
        self._ck(self.lib.gatx_timer_start(self.ctx), "gatx_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.gatx_timer_stop(self.ctx, C.byref(ms)), "gatx_timer_stop")
        return ms.value

    def launch_count(self):
        return self.lib.gatx_launch_count(self.ctx)

    def set_slopes(self, attn_slope=0.01, act_slope=0.01):
        """LeakyReLU slopes of the attention score / the layer activation (the reference fixes both at 0.01)"""
        self.lib.gatx_set_slopes.argtypes = [C.c_void_p, C.c_float, C.c_float]
        self._ck(self.lib.gatx_set_slopes(self.ctx, attn_slope, act_slope), "gatx_set_slopes")

    def set_bias(self, on=True):
        """Learnable per-layer bias on the aggregate (call before the parameters are set)"""
        self.lib.gatx_set_bias.argtypes = [C.c_void_p, C.c_int32]
        self._ck(self.lib.gatx_set_bias(self.ctx, int(on)), "gatx_set_bias")

    def set_bias_values(self, layer, b):
        b = np.ascontiguousarray(b, np.float32)
        self.lib.gatx_set_bias_values.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        self._ck(self.lib.gatx_set_bias_values(self.ctx, layer, b.ctypes.data), "gatx_set_bias_values")

    def set_dropout(self, p, seed=0):
        """Inverted dropout on every layer's input in training forwards (Philox, reproducible); p = 0 switches it off"""
        self.lib.gatx_set_dropout.argtypes = [C.c_void_p, C.c_float, C.c_uint64]
        self._ck(self.lib.gatx_set_dropout(self.ctx, p, seed), "gatx_set_dropout")

    def set_attn_dropout(self, p, seed=0):
        """Dropout on the attention coefficients in training forwards (Philox per (edge, head)); p = 0 switches it off"""
        self.lib.gatx_set_attn_dropout.argtypes = [C.c_void_p, C.c_float, C.c_uint64]
        self._ck(self.lib.gatx_set_attn_dropout(self.ctx, p, seed), "gatx_set_attn_dropout")

    def set_cuda_graph(self, mode):
        """-1 auto, 0 eager launches, 1 replay forward + backward of train_epoch as one CUDA graph"""
        self.lib.gatx_set_cuda_graph.argtypes = [C.c_void_p, C.c_int32]
        self._ck(self.lib.gatx_set_cuda_graph(self.ctx, mode), "gatx_set_cuda_graph")

    def cuda_graph_active(self):
        self.lib.gatx_cuda_graph_active.argtypes = [C.c_void_p]
        return bool(self.lib.gatx_cuda_graph_active(self.ctx))

    def edge_bytes(self, layer):
        f, b = C.c_double(), C.c_double()
        self._ck(self.lib.gatx_edge_bytes(self.ctx, layer, C.byref(f), C.byref(b)), "gatx_edge_bytes")
        return f.value, b.value

    def get_state(self):
        self.lib.gatx_state_size.restype = C.c_int64
        self.lib.gatx_state_size.argtypes = [C.c_void_p]
        n = self.lib.gatx_state_size(self.ctx)
        if n < 0:
            raise GatxError("gatx_state_size failed")
        out = np.empty(n, np.float32)
        self.lib.gatx_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self._ck(self.lib.gatx_get_state(self.ctx, out.ctypes.data, out.nbytes), "gatx_get_state")
        return out

    def set_state(self, state):
        state = np.ascontiguousarray(state, np.float32)
        self.lib.gatx_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self._ck(self.lib.gatx_set_state(self.ctx, state.ctypes.data, state.nbytes), "gatx_set_state")

    def set_train_mask(self, mask):
        """mask: uint8 [N] global (1 = node counts towards loss / accuracy / gradients) or None."""
        self.lib.gatx_set_train_mask.argtypes = [C.c_void_p, C.c_void_p]
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._ck(self.lib.gatx_set_train_mask(self.ctx, None if m is None else m.ctypes.data), "gatx_set_train_mask")

    def evaluate(self, mask=None):
        """Forward only; (avg_loss, accuracy) over the nodes of mask (None = all)."""
        self.lib.gatx_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lo, ac = C.c_float(), C.c_float()
        self._ck(self.lib.gatx_evaluate(self.ctx, None if m is None else m.ctypes.data, C.byref(lo), C.byref(ac)),
                 "gatx_evaluate")
        return lo.value, ac.value

    def peer_export(self):
        """This rank's GATX_PEER_INFO_BYTES blob (IPC handles of P_l / gP_l) for the NVLink halo exchange."""
        buf = (C.c_char * PEER_INFO_BYTES)()
        self.lib.gatx_peer_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self._ck(self.lib.gatx_peer_export(self.ctx, buf, PEER_INFO_BYTES), "gatx_peer_export")
        return bytes(buf)

    def peer_import(self, blobs):
        """blobs: list of every rank's peer_export() in rank order."""
        data = b"".join(blobs)
        self.lib.gatx_peer_import.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        self._ck(self.lib.gatx_peer_import(self.ctx, data, len(data)), "gatx_peer_import")

    def peer_disable(self):
        """Back to the NCCL collectives (call on EVERY rank when any rank failed to export / import)."""
        self.lib.gatx_peer_disable.argtypes = [C.c_void_p]
        self._ck(self.lib.gatx_peer_disable(self.ctx), "gatx_peer_disable")

    def halo_rows(self):
        self.lib.gatx_halo_rows.restype = C.c_int64
        self.lib.gatx_halo_rows.argtypes = [C.c_void_p]
        return int(self.lib.gatx_halo_rows(self.ctx))

    def halo_stats(self):
        """NVLink bytes / busy milliseconds of the exchange kernels in the last epoch (timing enabled)."""
        buf = (C.c_double * 4)()
        self.lib.gatx_halo_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self.lib.gatx_halo_stats(self.ctx, buf), "gatx_halo_stats")
        return dict(push_bytes=buf[0], push_ms=buf[1], pull_bytes=buf[2], pull_ms=buf[3])

    def halo_active(self):
        self.lib.gatx_halo_active.argtypes = [C.c_void_p]
        return bool(self.lib.gatx_halo_active(self.ctx))

    def comm_init(self, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.gatx_comm_init(self.ctx, buf), "gatx_comm_init")

"""ctypes binding of oracle/liborc.so -- TEST INFRASTRUCTURE ONLY (see gatv2_oracle.c header).

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f = C.POINTER(C.c_float)
c_i = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "gatv2_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_model_create.restype = C.c_void_p
        _LIB.orc_model_tensor.restype = c_f
        _LIB.orc_model_w_size.restype = C.c_int64
        _LIB.orc_model_loss.restype = C.c_double
        _LIB.orc_loss_acc.restype = C.c_double
        _LIB.orc_clip_grad_norm.restype = C.c_float
    return _LIB


def fp(a):
    return None if a is None else a.ctypes.data_as(c_f)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_i)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------- graph prep
def csr_to_coo(row_ptr, col_idx):
    N, E = len(row_ptr) - 1, len(col_idx)
    src, dst = np.empty(E, np.int32), np.empty(E, np.int32)
    lib().orc_csr_to_coo(N, ip(row_ptr), ip(col_idx), ip(src), ip(dst))
    return src, dst


def max_degree(row_ptr):
    return lib().orc_max_degree(ip(row_ptr), len(row_ptr) - 1)


def num_classes(labels):
    return lib().orc_num_classes(ip(labels), len(labels))


def csc_build(row_ptr, col_idx):
    N, E = len(row_ptr) - 1, len(col_idx)
    ptr, dst, eid = np.empty(N + 1, np.int32), np.empty(E, np.int32), np.empty(E, np.int32)
    lib().orc_csc_build(N, C.c_int64(E), ip(row_ptr), ip(col_idx), ip(ptr), ip(dst), ip(eid))
    return ptr, dst, eid


def partition_rows(row_ptr, R):
    b = np.empty(R + 1, np.int32)
    lib().orc_partition_rows(len(row_ptr) - 1, ip(row_ptr), R, ip(b))
    return b


# ------------------------------------------------------------- layer-level ops
def halo_rows(row_ptr, col_idx, R):
    """Per rank: number of (own source row, other rank) pairs where that rank's edge slice gathers the row -- the
    rows a rank sends per layer in the halo exchange (new work; bit-exact target for
    gatx_halo_rows).  Independent numpy formulation: unique (rank-of-edge, source) pairs."""
    b = partition_rows(row_ptr, R).astype(np.int64)
    row_ptr = np.asarray(row_ptr, np.int64)
    col_idx = np.asarray(col_idx, np.int64)
    N = len(row_ptr) - 1
    dst = np.repeat(np.arange(N, dtype=np.int64), np.diff(row_ptr))
    edge_rank = np.searchsorted(b[1:], dst, side="right")      # rank that owns the destination row
    pairs = np.unique(edge_rank * N + col_idx)
    ref_rank, src = pairs // N, pairs % N
    owner = np.searchsorted(b[1:], src, side="right")
    keep = owner != ref_rank
    return np.bincount(owner[keep], minlength=R).astype(np.int64)


def project(X, W, F):
    N, I = X.shape
    Pl, Pr = np.empty((N, F), np.float32), np.empty((N, F), np.float32)
    lib().orc_project(N, I, F, fp(X), fp(W), fp(Pl), fp(Pr))
    return Pl, Pr


def layer_forward(row_ptr, col_idx, H, D, Pl, Pr, a, is_last):
    N, E = len(row_ptr) - 1, len(col_idx)
    out = dict(
        score=np.zeros((H, E), np.float32), alpha=np.zeros((H, E), np.float32),
        mx=np.zeros((H, N), np.float32), sm=np.zeros((H, N), np.float32),
        hpre=np.zeros((N, H * D), np.float32),
        Hout=np.zeros((N, D if is_last else H * D), np.float32))
    lib().orc_layer_forward(N, ip(row_ptr), ip(col_idx), H, D, fp(Pl), fp(Pr), fp(a), int(is_last),
                            fp(out["score"]), fp(out["alpha"]), fp(out["mx"]), fp(out["sm"]),
                            fp(out["hpre"]), fp(out["Hout"]))
    return out


def layer_backward(row_ptr, col_idx, H, D, X, W, a, Pl, Pr, alpha, g_h, want_gx=True):
    N, E = len(row_ptr) - 1, len(col_idx)
    I = X.shape[1]
    F = H * D
    out = dict(gW=np.zeros((F, 2 * I), np.float32), ga=np.zeros(F, np.float32),
               gX=np.zeros((N, I), np.float32) if want_gx else None,
               galpha=np.zeros((H, E), np.float32), ge=np.zeros((H, E), np.float32),
               gPl=np.zeros((N, F), np.float32), gPr=np.zeros((N, F), np.float32))
    lib().orc_layer_backward(N, ip(row_ptr), ip(col_idx), H, D, I, fp(X), fp(W), fp(a), fp(Pl),
                             fp(Pr), fp(alpha), fp(g_h), fp(out["gW"]), fp(out["ga"]),
                             fp(out["gX"]), fp(out["galpha"]), fp(out["ge"]), fp(out["gPl"]),
                             fp(out["gPr"]))
    return out


def head_forward(Wo, HL):
    N, DL = HL.shape
    Cc = Wo.shape[0]
    z, y = np.empty((N, Cc), np.float32), np.empty((N, Cc), np.float32)
    lib().orc_head_forward(N, Cc, DL, fp(Wo), fp(HL), fp(z), fp(y))
    return z, y


def loss_acc(y, labels):
    N, Cc = y.shape
    losses, pred, corr = np.empty(N, np.float32), np.empty(N, np.int32), np.empty(N, np.int32)
    avg, acc = C.c_float(), C.c_float()
    total = lib().orc_loss_acc(N, Cc, fp(y), ip(labels), fp(losses), ip(pred), ip(corr),
                               C.byref(avg), C.byref(acc))
    return dict(total=total, avg=avg.value, acc=acc.value, losses=losses, pred=pred, correct=corr)


def output_grads(y, labels, hpre_last, HL, Wo, Hl):
    N, Cc = y.shape
    DL = HL.shape[1]
    gWo, g_h = np.zeros((Cc, DL), np.float32), np.zeros((N, Hl * DL), np.float32)
    lib().orc_output_grads(N, Cc, DL, Hl, fp(y), ip(labels), fp(hpre_last), fp(HL), fp(Wo), fp(gWo),
                           fp(g_h))
    return gWo, g_h


def lit_layer_forward(row_ptr, col_idx, H, D, X, W, a, is_last):
    N, E = len(row_ptr) - 1, len(col_idx)
    I = X.shape[1]
    out = dict(score=np.zeros((H, E), np.float32), alpha=np.zeros((H, E), np.float32),
               hpre=np.zeros((N, H * D), np.float32),
               Hout=np.zeros((N, D if is_last else H * D), np.float32))
    lib().orc_lit_layer_forward(N, ip(row_ptr), ip(col_idx), H, D, I, fp(X), fp(W), fp(a),
                                int(is_last), fp(out["score"]), fp(out["alpha"]), fp(out["hpre"]),
                                fp(out["Hout"]))
    return out


def lit_layer_backward(row_ptr, col_idx, H, D, X, W, a, alpha, g_h, want_gx=True):
    N, E = len(row_ptr) - 1, len(col_idx)
    I = X.shape[1]
    F = H * D
    out = dict(gW=np.zeros((F, 2 * I), np.float32), ga=np.zeros(F, np.float32),
               gX=np.zeros((N, I), np.float32) if want_gx else None,
               galpha=np.zeros((H, E), np.float32), ge=np.zeros((H, E), np.float32))
    lib().orc_lit_layer_backward(N, ip(row_ptr), ip(col_idx), H, D, I, fp(X), fp(W), fp(a),
                                 fp(alpha), fp(g_h), fp(out["gW"]), fp(out["ga"]), fp(out["gX"]),
                                 fp(out["galpha"]), fp(out["ge"]))
    return out


# tensor ids (identical to GATX_T_* in include/gatx.h)
T_W, T_A, T_WO, T_GW, T_GA, T_GWO, T_PL, T_PR, T_SCORE, T_ALPHA, T_HPRE, T_HOUT, T_Y, T_GH, T_Z = range(15)
T_B, T_GB = 26, 27  # bias extension (same ids as include/gatx.h)


class Model:
    """Whole-model oracle: forward, loss, backward, step exactly in EB's epoch order."""

    def __init__(self, heads, outdims, row_ptr, col_idx, X, labels, num_classes=None,
                 optimizer="sgd", clip=False, lr=1e-4, beta1=0.9, beta2=0.999):
        self.heads, self.outdims = list(heads), list(outdims)
        self.L = len(heads)
        self.row_ptr, self.col_idx = i32(row_ptr), i32(col_idx)
        self.X, self.labels = f32(X), i32(labels)
        self.N, self.I0 = self.X.shape
        self.E = len(self.col_idx)
        self.C = int(num_classes if num_classes is not None else self.labels.max() + 1)
        self.indims = [self.I0] + [h * d for h, d in zip(heads[:-1], outdims[:-1])]
        hs, od = i32(self.heads), i32(self.outdims)
        self._m = C.c_void_p(lib().orc_model_create(
            self.L, ip(hs), ip(od), self.N, C.c_int64(self.E), self.I0, self.C, ip(self.row_ptr),
            ip(self.col_idx), fp(self.X), ip(self.labels), 1 if optimizer == "adam" else 0,
            int(clip), C.c_float(lr), C.c_float(beta1), C.c_float(beta2)))

    def __del__(self):
        try:
            lib().orc_model_destroy(self._m)
        except Exception:
            pass

    def shape(self, which, l=0):
        H, D, I = self.heads[l], self.outdims[l], self.indims[l]
        F = H * D
        last = l == self.L - 1
        return {T_W: (F, 2 * I), T_GW: (F, 2 * I), T_A: (F,), T_GA: (F,), T_B: (F,), T_GB: (F,),
                T_WO: (self.C, self.outdims[-1]), T_GWO: (self.C, self.outdims[-1]),
                T_PL: (self.N, F), T_PR: (self.N, F), T_SCORE: (H, self.E), T_ALPHA: (H, self.E),
                T_HPRE: (self.N, F), T_HOUT: (self.N, D if last else F), T_Y: (self.N, self.C),
                T_Z: (self.N, self.C), T_GH: (self.N, F)}[which]

    def tensor(self, which, l=0):
        shp = self.shape(which, l)
        p = lib().orc_model_tensor(self._m, which, l)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shp)),)).reshape(shp)

    def set_params(self, l, W, a):
        lib().orc_model_set_params(self._m, l, fp(f32(W)), fp(f32(a)))

    def set_wo(self, Wo):
        lib().orc_model_set_wo(self._m, fp(f32(Wo)))

    def set_mask(self, mask):
        """mask: uint8 [N] (1 = node counts towards loss / accuracy / gradients) or None."""
        self._mask = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_model_set_mask(self._m, None if mask is None else self._mask.ctypes.data_as(C.c_void_p))

    def set_bias(self, on=True):
        lib().orc_model_set_bias(self._m, int(on))

    def set_bias_values(self, l, b):
        lib().orc_model_set_bias_values(self._m, l, fp(f32(b)))

    def set_attn_dropout(self, p, seed):
        """Attention-coefficient dropout in training forwards (oracle only so far); resets the step counter."""
        lib().orc_model_set_attn_dropout(self._m, C.c_float(p), C.c_uint64(seed))

    def set_dropout(self, p, seed):
        """Dropout on every layer's input in training forwards; resets the step counter."""
        lib().orc_model_set_dropout(self._m, C.c_float(p), C.c_uint64(seed))

    def forward(self, train=True):
        (lib().orc_model_forward if train else lib().orc_model_forward_eval)(self._m)

    def loss(self):
        avg, acc = C.c_float(), C.c_float()
        pred = np.empty(self.N, np.int32)
        total = lib().orc_model_loss(self._m, C.byref(avg), C.byref(acc), ip(pred))
        return dict(total=total, avg=avg.value, acc=acc.value, pred=pred)

    def backward(self):
        lib().orc_model_backward(self._m)

    def step(self, t):
        lib().orc_model_step(self._m, t)

    def epoch(self, t):
        avg, acc = C.c_float(), C.c_float()
        lib().orc_model_epoch(self._m, t, C.byref(avg), C.byref(acc))
        return avg.value, acc.value


def set_slopes(attn=0.01, act=0.01):
    """LeakyReLU slopes of the attention score / layer activation (process-wide; the reference fixes 0.01 / 0.01)."""
    lib().orc_set_slopes(C.c_float(attn), C.c_float(act))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, out)
    return [int(v) for v in out]


def dropout(X, p, seed, layer, step, row0=0):
    """Inverted dropout of a [rows][cols] block (keep rule and scale documented at orc_dropout)."""
    X = f32(X)
    Y = np.empty_like(X)
    lib().orc_dropout(fp(X), fp(Y), X.shape[0], X.shape[1], row0, C.c_float(p), C.c_uint64(seed), layer,
                      C.c_int64(step))
    return Y


def attn_dropout_scale(H, E, p, seed, layer, step):
    out = np.empty((H, E), np.float32)
    lib().orc_attn_dropout_scale(fp(out), H, C.c_int64(E), C.c_float(p), C.c_uint64(seed), layer, C.c_int64(step))
    return out


def num_threads():
    return lib().orc_num_threads()


def row_blocks(row_ptr, R, K):
    """Blocks of the pipelined multi-GPU exchange (new work; bit-exact target for gatx_row_blocks): rank p's own rows cut
    into K edge-balanced blocks, out[p][k] = first own row whose edge offset (relative to the rank's first edge) reaches
    floor(k E_p / K).  Independent numpy formulation (searchsorted on the rebased row_ptr)."""
    b = partition_rows(row_ptr, R).astype(np.int64)
    rp = np.asarray(row_ptr, np.int64)
    out = np.empty((R, K + 1), np.int32)
    for p in range(R):
        base, ep = rp[b[p]], rp[b[p + 1]] - rp[b[p]]
        targets = ep * np.arange(K + 1, dtype=np.int64) // K
        rel = rp[b[p]:b[p + 1] + 1] - base
        out[p] = b[p] + np.searchsorted(rel[:-1] if len(rel) > 1 else rel, targets, side="left")
        out[p][0], out[p][K] = b[p], b[p + 1]
    return out

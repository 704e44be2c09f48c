"""Independent dense/autograd restatement of the GATv2 epoch math in PyTorch float64.

Used only to pin the C oracle (tests/test_oracle.py): forward values are compared directly and
the oracle's hand-derived gradients are compared with autograd of the SUMMED cross-entropy
(GATv2_edge_based.cu:572 uses y - onehot with no 1/N).
"""
import numpy as np
import torch

SLOPE = 0.01


def lrelu(x, slope=SLOPE):
    return torch.where(x > 0, x, slope * x)


def forward(Ws, As, Wo, X, row_ptr, col_idx, heads, outdims, labels, dtype=torch.float64, mask=None,
            attn_slope=SLOPE, act_slope=SLOPE, in_scale=None, biases=None, alpha_scale=None):
    """in_scale: optional per-layer [N][I_l] arrays multiplied into each layer's input (a dropout mask / (1 - p))."""
    N = len(row_ptr) - 1
    deg = np.diff(row_ptr)
    src = torch.as_tensor(np.asarray(col_idx), dtype=torch.long)
    dst = torch.as_tensor(np.repeat(np.arange(N), deg), dtype=torch.long)
    E = len(src)
    x = torch.as_tensor(X, dtype=dtype)
    out = dict(Pl=[], Pr=[], score=[], alpha=[], hpre=[], Hout=[])
    L = len(heads)
    for l, (H, D) in enumerate(zip(heads, outdims)):
        I = x.shape[1]
        W, a = Ws[l], As[l]
        if in_scale is not None:
            x = x * torch.as_tensor(np.asarray(in_scale[l], np.float64), dtype=dtype)
        Pl = x @ W[:, :I].T
        Pr = x @ W[:, I:].T
        s = (Pl[src] + Pr[dst]).view(E, H, D)
        score = (lrelu(s, attn_slope) * a.view(1, H, D)).sum(-1)  # [E, H]
        m = torch.full((N, H), -1e9, dtype=dtype).scatter_reduce(
            0, dst[:, None].expand(E, H), score.detach(), "amax", include_self=True)
        ex = torch.exp(score - m[dst])
        ssum = torch.zeros((N, H), dtype=dtype).index_add(0, dst, ex)
        alpha = ex / (ssum[dst] + 1e-8)
        agg = alpha if alpha_scale is None else alpha * torch.as_tensor(np.asarray(alpha_scale[l], np.float64).T, dtype=dtype)
        h = torch.zeros((N, H, D), dtype=dtype).index_add(0, dst, agg[:, :, None] * Pl[src].view(E, H, D))
        if biases is not None:
            h = h + biases[l].view(1, H, D)
        if l == L - 1:
            Hout = lrelu(h, act_slope).mean(1)
        else:
            Hout = lrelu(h, act_slope).reshape(N, H * D)
        for k, v in (("Pl", Pl), ("Pr", Pr), ("score", score.T), ("alpha", alpha.T),
                     ("hpre", h.reshape(N, H * D)), ("Hout", Hout)):
            out[k].append(v)
        x = Hout
    z = x @ Wo.T
    zm = z.max(dim=1, keepdim=True).values.detach()
    ez = torch.exp(z - zm)
    y = ez / (ez.sum(1, keepdim=True) + 1e-8)
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.long)
    p = y[torch.arange(N), lab]
    losses = -torch.log(torch.clamp(p, min=1e-12))
    if mask is not None:  # extension: only counted nodes enter the (summed) loss
        losses = losses * torch.as_tensor(np.asarray(mask) != 0, dtype=dtype)
    out.update(z=z, y=y, losses=losses, loss_sum=losses.sum(), pred=y.argmax(1))
    return out


def forward_backward(Ws, As, Wo, X, row_ptr, col_idx, heads, outdims, labels, mask=None, **kw):
    tW = [torch.tensor(np.asarray(w, np.float64), requires_grad=True) for w in Ws]
    tA = [torch.tensor(np.asarray(a, np.float64), requires_grad=True) for a in As]
    tWo = torch.tensor(np.asarray(Wo, np.float64), requires_grad=True)
    tB = None
    if kw.get("biases") is not None:
        tB = [torch.tensor(np.asarray(b, np.float64), requires_grad=True) for b in kw["biases"]]
        kw = dict(kw, biases=tB)
    out = forward(tW, tA, tWo, np.asarray(X, np.float64), row_ptr, col_idx, heads, outdims, labels, mask=mask, **kw)
    out["loss_sum"].backward()
    grads = dict(gW=[w.grad.numpy() for w in tW], ga=[a.grad.numpy() for a in tA], gWo=tWo.grad.numpy())
    if tB is not None:
        grads["gb"] = [b.grad.numpy() for b in tB]
    vals = {k: ([t.detach().numpy() for t in v] if isinstance(v, list) else v.detach().numpy())
            for k, v in out.items()}
    return vals, grads

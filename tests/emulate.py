"""CPU interpreter of the implicit-GEMM descriptors that cstp_b200.ops hands to the CUDA kernels.

It evaluates exactly what the kernels are specified to compute (tap lists over stride-parity views with zero
out-of-range reads, tile-space scatter with offsets/strides) in fp32 torch, so the host-side geometry can be
validated against torch's convolution functions without a GPU.  Test infrastructure only.
"""
import torch

from cstp_b200 import ops


def _gather(x, parity, stride, d, space):
    """x: (N,T,H,W,C) -> values of the parity view at tile-space coords + d, zero outside the view."""
    N, T, H, W, C = x.shape
    (rt, rh, rw), (st, sh, sw) = parity, stride
    sub = x[:, rt::st, rh::sh, rw::sw, :]
    Wt, Ht, Tt, Nt = space
    dw, dh, dt = d
    out = torch.zeros(Nt, Tt, Ht, Wt, C, dtype=x.dtype)
    ts = torch.arange(Tt) + dt
    hs = torch.arange(Ht) + dh
    ws = torch.arange(Wt) + dw
    tv = (ts >= 0) & (ts < sub.shape[1])
    hv = (hs >= 0) & (hs < sub.shape[2])
    wv = (ws >= 0) & (ws < sub.shape[3])
    if tv.any() and hv.any() and wv.any():
        blk = sub[:, ts[tv]][:, :, hs[hv]][:, :, :, ws[wv]]
        ti = torch.nonzero(tv).flatten()
        hi = torch.nonzero(hv).flatten()
        wi = torch.nonzero(wv).flatten()
        out[:, ti[:, None, None], hi[None, :, None], wi[None, None, :]] = blk
    return out


def emulate_fwd(x, w, geom):
    """x (N,T,H,W,Ci) fp32, w (Co,Ci,kt,kh,kw) -> (N,To,Ho,Wo,Co) via the forward tap list."""
    N, T, H, W, Ci = x.shape
    To, Ho, Wo = geom.out_dims(T, H, W)
    maps, taps = ops.fwd_taps(geom)
    wk = w.reshape(w.shape[0], Ci, -1)
    out = torch.zeros(N, To, Ho, Wo, w.shape[0])
    for (m, dw, dh, dt, ti) in taps:
        a = _gather(x, maps[m], geom.stride, (dw, dh, dt), (Wo, Ho, To, N))
        out += a @ wk[:, :, ti].t()
    return out


def emulate_dgrad(g, w, geom, dx_shape):
    """g (N,To,Ho,Wo,Co) -> dx (N,T,H,W,Ci) via the per-parity-class tap lists; returns (dx, covers_all)."""
    N, T, H, W, Ci = dx_shape
    wk = w.reshape(w.shape[0], Ci, -1)
    dx = torch.zeros(dx_shape)
    flat = dx.view(-1)
    covers = True
    for cl in ops.dgrad_classes(dx_shape, geom):
        if not cl["taps"]:
            covers = False
            continue
        Wt, Ht, Tt, Nt = cl["space"]
        acc = torch.zeros(Nt, Tt, Ht, Wt, Ci)
        for (dw, dh, dt, ti) in cl["taps"]:
            a = _gather(g, (0, 0, 0), (1, 1, 1), (dw, dh, dt), cl["space"])
            acc += a @ wk[:, :, ti]
        osw, osh, ost, osn = cl["ostrides"]
        idx = (cl["off"] + torch.arange(Nt)[:, None, None, None] * osn + torch.arange(Tt)[None, :, None, None] * ost
               + torch.arange(Ht)[None, None, :, None] * osh + torch.arange(Wt)[None, None, None, :] * osw)
        idx = idx[..., None] + torch.arange(Ci)
        flat[idx.reshape(-1)] = acc.reshape(-1)
    return dx, covers


def emulate_wgrad(x, g, geom, cout, cin):
    """dW (cout,cin,kt,kh,kw) from the wgrad chunk list (tap shift applied to x, positions summed)."""
    N, T, H, W, Ci = x.shape
    _, To, Ho, Wo, Co = g.shape
    maps, taps = ops.fwd_taps(geom)
    dw_ = torch.zeros(cout, cin, geom.taps)
    for (m, dw, dh, dt, ti) in taps:
        a = _gather(x, maps[m], geom.stride, (dw, dh, dt), (Wo, Ho, To, N))
        dw_[:, :, ti] = (g.reshape(-1, Co).t() @ a.reshape(-1, Ci))[:cout, :cin]
    return dw_.view(cout, cin, *geom.kernel)

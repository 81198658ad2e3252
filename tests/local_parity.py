"""Per-layer ("local") parity of the step engine against fp32 torch ops -- TEST INFRASTRUCTURE.

Every layer of the recorded engine step is re-evaluated on CPU in fp32 from the engine's OWN inputs of that layer
(its bf16 activations / gradients, upcast) with the reference's operators (F.conv3d, F.batch_norm, F.relu, F.linear and
their autograd), so that each comparison measures one layer's arithmetic (bf16 operand/weight/output rounding + fp32
accumulation order) and not the drift accumulated through the 24 layers in front of it.  End-to-end drift against
the fp32 oracle is measured separately (tests/test_gpu_step.py, DESIGN.md "Numerics").

Works on any engine object that was built with record=True and has run one train_step: the CUDA engine (GPU tests)
or the engine on top of tests/emulate_ops.py (CPU tests of the checker itself).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _f(t):
    return t.detach().float().cpu()


def _ncdhw(t, C):
    return _f(t)[..., :C].permute(0, 4, 1, 2, 3).contiguous()


def _bn_groups(x, gamma, beta, groups=2):
    """Training-mode BatchNorm with separate statistics per view (the reference pushes the views one by one)."""
    parts = x.chunk(groups, 0)
    return torch.cat([F.batch_norm(p, None, None, gamma, beta, True, 0.0, BN_EPS) for p in parts], 0)


def check_conv_units(eng, params, which="online"):
    """params: name -> fp32 CPU tensor of the weights the step USED (pre-step online, post-EMA target).
    Returns {tag: {metric: rel_err}} for every conv+BN unit of the chosen network, plus dx checks keyed 'dx:<tag>'."""
    out: dict = {}
    units = [u for u in eng.units if u["tag"].startswith(which + ".")]
    by_raw = {u["raw"].data_ptr(): u for u in units}
    dx_ref: dict = {}        # data_ptr of an activation -> accumulated reference gradient (NCDHW fp32)
    dx_shape: dict = {}
    for u in units:
        tag, geom, cin, cout = u["tag"], u["geom"], u["cin"], u["cout"]
        W = params[u["wname"]].clone().requires_grad_(u["grads"])
        e: dict = {}
        # ---------------------------------------------------------------- convolution
        if u["stem"]:
            # the engine's operand is the packed stem row pairs P[n][t][h2][wo][hpar*32 + kw*3 + c] =
            # x[n][c][t][2*h2 + hpar][2*wo + kw - 3]: the bf16 clip is read back from the kw = 3 (even columns) and
            # kw = 4 (odd columns) taps
            P = _f(u["x"])
            Nn, Tt, H2, Wo_, _ = P.shape
            P = P.reshape(Nn, Tt, H2, Wo_, 2, 32)
            xin = torch.stack([P[..., 9:12], P[..., 12:15]], 5)          # n, t, h2, wo, hpar, wpar, c
            xin = xin.permute(0, 6, 1, 2, 4, 3, 5).reshape(Nn, 3, Tt, 2 * H2, 2 * Wo_).contiguous()
            raw_r = F.conv3d(xin, W, None, (1, 2, 2), (0, 3, 3))
            raw_e5 = _ncdhw(u["raw"], cout)
            e["conv"] = rel(raw_e5, raw_r)
        else:
            # (x_act: with the BatchNorm apply fused into this conv's operand path, u["x"] is the producer's raw output and
            # x_act the activation the engine materialised for the tests -- the values the tensor cores consumed)
            xin = _ncdhw(u["x_act"], cin).requires_grad_(u["grads"] and not u["skip_dgrad"])
            raw_r = F.conv3d(xin, W, None, geom.stride, geom.pad)
            raw_e5 = _ncdhw(u["raw"], cout)
            e["conv"] = rel(raw_e5, raw_r)
        # ---------------------------------------------------------------- BatchNorm (+ residual) (+ ReLU)
        gamma = params[u["bnname"] + ".weight"].clone().requires_grad_(u["grads"])
        beta = params[u["bnname"] + ".bias"].clone().requires_grad_(u["grads"])
        leaf = raw_e5.clone().requires_grad_(u["grads"])
        views = getattr(eng, "VIEWS", 2)
        y = _bn_groups(leaf, gamma, beta, views)
        if u["res"] is not None:
            if u["res_site"] is not None:
                ds = by_raw[u["res"].data_ptr()]
                r = _bn_groups(_ncdhw(u["res"], cout), params[ds["bnname"] + ".weight"], params[ds["bnname"] + ".bias"],
                               views)
            else:
                r = _ncdhw(u["res_act"], cout)
            y = y + r
        act_e = None
        if u["act"] is not None:
            act_e = _ncdhw(u["act"], cout)
            e["bn_act"] = rel(act_e, F.relu(y) if u["relu"] else y)
        if u["relu"]:
            # ReLU backward is defined by the sign of the FORWARD OUTPUT (torch: grad * (result > 0)); the engine's own
            # output is used as that result so that exact ties at the kink (fmaf vs mul+add, +-1e-10) cannot flip a mask.
            mask = (act_e > 0).to(y.dtype)
            y = y * mask
        if not u["grads"]:
            out[tag] = e
            continue
        # ---------------------------------------------------------------- BatchNorm / ReLU backward
        d_out = eng.named[tag + ".d_out"]
        d5 = _ncdhw(d_out, cout)
        y.backward(d5)
        g_e = eng.named[tag + ".g"]
        g_e5 = _ncdhw(g_e, cout)
        e["bn_bwd_dx"] = rel(g_e5, leaf.grad)
        e["bn_dgamma"] = rel(eng.train.view(u["bnname"] + ".weight", eng.grad), gamma.grad)
        e["bn_dbeta"] = rel(eng.train.view(u["bnname"] + ".bias", eng.grad), beta.grad)
        # padded channels of the gradient must stay exactly zero (they feed wgrad / dgrad as K columns)
        e["pad_zero"] = float(_f(g_e)[..., cout:].abs().max()) if g_e.shape[-1] > cout else 0.0
        # ---------------------------------------------------------------- conv backward from the ENGINE's d(raw)
        raw_r.backward(g_e5)
        e["wgrad"] = rel(eng.train.view(u["wname"], eng.grad), W.grad.reshape(params[u["wname"]].shape))
        if not u["skip_dgrad"]:
            k = u["x"].data_ptr()
            dx_ref[k] = dx_ref.get(k, 0) + xin.grad
            dx_shape[k] = (u["x"], cin)
        if u["res"] is not None and u["res_site"] is None:
            # identity shortcut: the masked upstream gradient flows straight into d(block input)
            k = u["res"].data_ptr()
            dz = d5 * (_ncdhw(u["act"], cout) > 0)
            dx_ref[k] = dx_ref.get(k, 0) + dz
            dx_shape[k] = (u["res"], cout)
        out[tag] = e
    for k, ref in dx_ref.items():
        t, C = dx_shape[k]
        name = next(n for n, v in eng.named.items() if torch.is_tensor(v) and v.data_ptr() == k and
                    n.endswith((".act", ".out", ".raw")))
        out["dx:" + name] = {"dgrad": rel(_ncdhw(eng._dbuf(t), C), ref)}
    return out


def check_mlp(eng, params, tag, pre, cin, hidden, cout, groups, names=("0", "1", "3")):
    """Linear -> BatchNorm1d -> ReLU -> Linear head: forward and (when recorded) backward, layer by layer."""
    a, b, c = names
    n = eng.named
    e: dict = {}
    x = _f(n[tag + ".x"])[:, :cin]
    W0, b0 = params[f"{pre}.{a}.weight"], params[f"{pre}.{a}.bias"]
    W3, b3 = params[f"{pre}.{c}.weight"], params[f"{pre}.{c}.bias"]
    gamma = params[f"{pre}.{b}.weight"].clone().requires_grad_(True)
    beta = params[f"{pre}.{b}.bias"].clone().requires_grad_(True)
    raw_e = _f(n[tag + ".raw"])[:, :hidden]
    e["linear0"] = rel(raw_e, F.linear(x, W0, b0))
    leaf = raw_e.clone().requires_grad_(True)
    h_e = _f(n[tag + ".h"])[:, :hidden]
    h = _bn_groups(leaf, gamma, beta, groups)
    e["bn_act"] = rel(h_e, F.relu(h))
    h = h * (h_e > 0).to(h.dtype)         # ReLU backward masks on the forward output (see check_conv_units)
    e["linear3"] = rel(_f(n[tag + ".out"])[:, :cout], F.linear(h_e, W3, b3))
    if tag + ".g_out" not in n:
        return e
    g_out = _f(n[tag + ".g_out"])[:, :cout]
    e["wgrad3"] = rel(eng.train.view(f"{pre}.{c}.weight", eng.grad), g_out.t() @ h_e)
    e["dbias3"] = rel(eng.train.view(f"{pre}.{c}.bias", eng.grad), g_out.sum(0))
    d_h = _f(n[tag + ".d_h"])[:, :hidden]
    e["dgrad3"] = rel(d_h, g_out @ W3)
    h.backward(d_h)
    g_h = _f(n[tag + ".g_h"])[:, :hidden]
    e["bn_bwd_dx"] = rel(g_h, leaf.grad)
    e["bn_dgamma"] = rel(eng.train.view(f"{pre}.{b}.weight", eng.grad), gamma.grad)
    e["bn_dbeta"] = rel(eng.train.view(f"{pre}.{b}.bias", eng.grad), beta.grad)
    e["wgrad0"] = rel(eng.train.view(f"{pre}.{a}.weight", eng.grad), g_h.t() @ x)
    return e


MLPS = [("online.project", "online_net.project.net", 512, 4096, 512, 2),
        ("predictor", "predictor.net", 512, 4096, 512, 2),
        ("overlap_spa", "overlap_spa", 1024, 1024, 5, 1),
        ("overlap_tem", "overlap_tem", 1024, 1024, 5, 1),
        ("pb_cls", "pb_cls", 512, 512, 5, 2),
        ("rotate_cls", "rotate_cls", 512, 512, 5, 2)]


def check_all(eng, online_params, target_params):
    res = check_conv_units(eng, online_params, "online")
    res.update(check_conv_units(eng, target_params, "target"))
    for tag, pre, cin, hidden, cout, groups in MLPS:
        res["mlp:" + tag] = check_mlp(eng, online_params, tag, pre, cin, hidden, cout, groups)
    res["mlp:target.project"] = check_mlp(eng, target_params, "target.project", "target_net.project.net", 512, 4096,
                                          512, 2)
    return res


def worst(res, metrics=None):
    """(max error, tag, metric) over the result table, optionally restricted to some metrics."""
    items = [(v, tag, m) for tag, e in res.items() for m, v in e.items() if m != "pad_zero" and
             (metrics is None or m in metrics)]
    return max(items) if items else (0.0, "", "")

"""CPU oracle for the CSTP `r21d_byol` pretraining hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 restatement (torch CPU ops, no nn.Module, parameters passed as a name->tensor dict keyed by the
reference's state_dict names) of what the reference computes on this path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import it; the product path (cstp_b200/) never does.

Pinning: the reference ships no tests, golden vectors or fixtures, so parity is pinned
against the reference ITSELF, executed in the build container: oracle/make_golden.py imports the unmodified
reference modules from /root/reference, runs the seeded protocol of SURVEY.md A.2 and stores the outputs in
tests/golden/; tests/test_oracle_golden.py checks this restatement against those vectors.

Every function cites the reference lines it follows (paths relative to the reference repo root).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------- architecture
def intermed_channels(cin: int, cout: int, k: tuple[int, int, int]) -> int:
    """models/pace/r21d_byol.py:74-76 -- number of channels between the spatial and temporal halves."""
    kt, kh, kw = k
    return int(math.floor((kt * kh * kw * cin * cout) / (kh * kw * cin + kt * cout)))


def backbone_layout():
    """The 12 SpatioTemporalConv instances of R2Plus1DNet((1,1,1,1)) -- r21d_byol.py:184-210,100-139.
    Returns [(prefix, cin, cout, kernel, stride, pad)] in registration order."""
    L = [("conv1", 3, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3))]
    cin = 64
    for stage, cout, down in (("conv2", 64, False), ("conv3", 128, True), ("conv4", 256, True), ("conv5", 512, True)):
        b = f"{stage}.block1"
        if down:
            L.append((b + ".downsampleconv", cin, cout, (1, 1, 1), (2, 2, 2), (0, 0, 0)))
            L.append((b + ".conv1", cin, cout, (3, 3, 3), (2, 2, 2), (1, 1, 1)))
        else:
            L.append((b + ".conv1", cin, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1)))
        L.append((b + ".conv2", cout, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1)))
        cin = cout
    return L


class Tape:
    """Records named intermediate tensors (and keeps their grads) for per-layer parity checks."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self.acts: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def rec(self, name: str, t: torch.Tensor) -> torch.Tensor:
        if self.enabled:
            if t.requires_grad:
                t.retain_grad()
            self.acts[name] = t
        return t


def _bn(x, P, name, buffers_out, tape, tag):
    """nn.BatchNorm3d / BatchNorm1d in training mode (r21d_byol.py:83,126,133,138,199,237,251,277-290):
    batch statistics, affine, running-stat update with momentum 0.1 and unbiased variance."""
    rm = P[name + ".running_mean"].clone()
    rv = P[name + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, P[name + ".weight"], P[name + ".bias"], True, BN_MOMENTUM, BN_EPS)
    if buffers_out is not None:
        # a second call on the same module (second view) continues from the first call's buffers
        buffers_out[name + ".running_mean"] = rm
        buffers_out[name + ".running_var"] = rv
        P = P  # (caller re-reads buffers_out through _params_view)
    return tape.rec(tag, y)


class _ParamView(dict):
    """name->tensor lookup that prefers updated BN buffers over the initial ones."""

    def __init__(self, base, updated):
        super().__init__()
        self.base, self.updated = base, updated

    def __getitem__(self, k):
        return self.updated[k] if k in self.updated else self.base[k]


def st_conv(x, P, pre, kernel, stride, pad, bufs, tape, tag):
    """SpatioTemporalConv.forward -- r21d_byol.py:94-97 with the decomposition of :58-70."""
    kt, kh, kw = kernel
    x = F.conv3d(x, P[pre + ".spatial_conv.weight"], None, (1, stride[1], stride[2]), (0, pad[1], pad[2]))
    tape.rec(tag + ".spatial_conv", x)
    x = F.relu(_bn(x, P, pre + ".bn", bufs, tape, tag + ".bn"))
    x = F.conv3d(x, P[pre + ".temporal_conv.weight"], None, (stride[0], 1, 1), (pad[0], 0, 0))
    return tape.rec(tag + ".temporal_conv", x)


def res_block(x, P, pre, cin, cout, down, bufs, tape, tag):
    """SpatioTemporalResBlock.forward -- r21d_byol.py:141-148."""
    s = (2, 2, 2) if down else (1, 1, 1)
    res = st_conv(x, P, pre + ".conv1", (3, 3, 3), s, (1, 1, 1), bufs, tape, tag + ".conv1")
    res = F.relu(_bn(res, P, pre + ".bn1", bufs, tape, tag + ".bn1"))
    res = st_conv(res, P, pre + ".conv2", (3, 3, 3), (1, 1, 1), (1, 1, 1), bufs, tape, tag + ".conv2")
    res = _bn(res, P, pre + ".bn2", bufs, tape, tag + ".bn2")
    if down:
        x = st_conv(x, P, pre + ".downsampleconv", (1, 1, 1), (2, 2, 2), (0, 0, 0), bufs, tape, tag + ".downsampleconv")
        x = _bn(x, P, pre + ".downsamplebn", bufs, tape, tag + ".downsamplebn")
    return tape.rec(tag + ".out", F.relu(x + res))


def mlp(x, P, pre, bufs, tape, tag, names=("0", "1", "3")):
    """Projector / Predictor / pretext heads: Linear -> BatchNorm1d -> ReLU -> Linear -- r21d_byol.py:232-257,276-291."""
    a, b, c = names
    x = F.linear(x, P[f"{pre}.{a}.weight"], P[f"{pre}.{a}.bias"])
    x = F.relu(_bn(x, P, f"{pre}.{b}", bufs, tape, tag + ".bn"))
    return tape.rec(tag + ".out", F.linear(x, P[f"{pre}.{c}.weight"], P[f"{pre}.{c}.bias"]))


def r2plus1d_net(x, P, pre, bufs, tape, tag, proj=True):
    """R2Plus1DNet.forward -- r21d_byol.py:215-229.  x: (B,3,T,H,W) fp32."""
    x = st_conv(x, P, pre + ".conv1", (3, 7, 7), (1, 2, 2), (1, 3, 3), bufs, tape, tag + ".conv1")
    x = tape.rec(tag + ".stem.out", F.relu(_bn(x, P, pre + ".bn1", bufs, tape, tag + ".bn1")))
    cin = 64
    for stage, cout, down in (("conv2", 64, False), ("conv3", 128, True), ("conv4", 256, True), ("conv5", 512, True)):
        x = res_block(x, P, f"{pre}.{stage}.block1", cin, cout, down, bufs, tape, f"{tag}.{stage}.block1")
        cin = cout
    feat = tape.rec(tag + ".feat", F.adaptive_avg_pool3d(x, 1).view(-1, 512))
    if not proj:
        return feat
    return feat, mlp(feat, P, pre + ".project.net", bufs, tape, tag + ".project")


def byol_loss_fn(x, y):
    """R21DBYOL._loss_fn -- r21d_byol.py:346-349."""
    x = F.normalize(x, dim=-1, p=2)
    y = F.normalize(y, dim=-1, p=2)
    return 2 - 2 * (x * y).sum(dim=-1)


def ema_update(target: dict, online: dict, momentum: float = 0.996) -> None:
    """R21DBYOL._update_target_net -- r21d_byol.py:331-337 (parameters only, in place on `target`)."""
    for k in list(target.keys()):
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            continue
        target[k] = target[k] * momentum + online[k] * (1. - momentum)


def loss_com_forward(state: dict, x1, x2, momentum=0.996, tape: Tape | None = None):
    """R21DBYOL.forward(x1, x2, o_type="loss_com") -- r21d_byol.py:357-382.

    `state` maps reference state_dict names (no DDP `module.` prefix) to fp32 tensors; trainable entries may
    require grad.  Returns (loss_byol, 6 logits, new_buffers, new_target_params)."""
    tape = tape or Tape(False)
    bufs: dict = {}
    P = _ParamView(state, bufs)
    f1, p1 = r2plus1d_net(x1, P, "online_net", bufs, tape, "online.v1")
    f2, p2 = r2plus1d_net(x2, P, "online_net", bufs, tape, "online.v2")
    q1 = mlp(p1, P, "predictor.net", bufs, tape, "predictor.v1")
    q2 = mlp(p2, P, "predictor.net", bufs, tape, "predictor.v2")
    with torch.no_grad():
        tgt = {k[len("target_net."):]: v.detach() for k, v in state.items() if k.startswith("target_net.")}
        onl = {k[len("online_net."):]: v.detach() for k, v in state.items() if k.startswith("online_net.")}
        ema_update(tgt, onl, momentum)
        new_target = {"target_net." + k: v for k, v in tgt.items()}
        PT = _ParamView({**state, **new_target}, bufs)
        _, t1 = r2plus1d_net(x1, PT, "target_net", bufs, tape, "target.v1")
        _, t2 = r2plus1d_net(x2, PT, "target_net", bufs, tape, "target.v2")
    loss = (byol_loss_fn(q1, t2) + byol_loss_fn(q2, t1)).mean()          # r21d_byol.py:351-355,382
    cat = torch.cat((f1, f2), dim=1)                                      # :374
    logits = (mlp(cat, P, "overlap_spa", bufs, tape, "overlap_spa"),      # :375-380
              mlp(cat, P, "overlap_tem", bufs, tape, "overlap_tem"),
              mlp(f1, P, "pb_cls", bufs, tape, "pb_cls.v1"),
              mlp(f2, P, "pb_cls", bufs, tape, "pb_cls.v2"),
              mlp(f1, P, "rotate_cls", bufs, tape, "rotate_cls.v1"),
              mlp(f2, P, "rotate_cls", bufs, tape, "rotate_cls.v2"))
    return loss, logits, bufs, new_target


def total_loss(loss_byol, logits, labels, loss_weight):
    """main_byol.py:62-73 -- six nn.CrossEntropyLoss() terms and the --loss_weight combination.
    labels = (spa, tem, pb, rot1, rot2) int64."""
    spa, tem, pb, r1, r2 = labels
    ce = [F.cross_entropy(logits[0], spa), F.cross_entropy(logits[1], tem), F.cross_entropy(logits[2], pb),
          F.cross_entropy(logits[3], pb), F.cross_entropy(logits[4], r1), F.cross_entropy(logits[5], r2)]
    w = loss_weight
    tot = w[0] * loss_byol + w[1] * ce[0] + w[2] * ce[1] + w[3] * ce[2] + w[3] * ce[3] + w[4] * ce[4] + w[4] * ce[5]
    return tot, ce


def clip_and_sgd(params: dict, grads: dict, mom: dict, lr, momentum=0.9, wd=5e-4, max_norm=18.0, clip=True):
    """main_byol.py:88-91 + optim.SGD(lr, momentum, weight_decay) of :229-232, restated on dicts (in place).
    Returns the total gradient norm before clipping."""
    names = [k for k in params if grads.get(k) is not None]
    total = torch.norm(torch.stack([torch.norm(grads[k].detach(), 2.0) for k in names]), 2.0)
    coef = 1.0
    if clip:
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for k in names:
        g = grads[k] * coef
        g = g.add(params[k], alpha=wd)
        if k not in mom:
            mom[k] = g.clone()
        else:
            mom[k].mul_(momentum).add_(g)
        params[k] = params[k] - lr * mom[k]
    return float(total)


def pretrain_step(state: dict, trainable: list[str], x1, x2, labels, loss_weight, lr, mom: dict, tape=None,
                  momentum_ema=0.996, clip=True):
    """One full step of main_byol.py:60-91 on `state` (updated in place).  Returns a dict of scalars/tensors."""
    for k in trainable:
        state[k] = state[k].detach().requires_grad_(True)
    loss_byol, logits, bufs, new_target = loss_com_forward(state, x1, x2, momentum_ema, tape)
    tot, ce = total_loss(loss_byol, logits, labels, loss_weight)
    grads_list = torch.autograd.grad(tot, [state[k] for k in trainable], allow_unused=True, retain_graph=tape is not None)
    if tape is not None and tape.enabled:
        # populate .grad of the recorded activations as well
        keep = [t for t in tape.acts.values() if t.requires_grad]
        agr = torch.autograd.grad(tot, keep, allow_unused=True)
        tape.act_grads = {n: g for (n, t), g in zip([(n, t) for n, t in tape.acts.items() if t.requires_grad], agr)}
    grads = dict(zip(trainable, grads_list))
    params = {k: state[k].detach() for k in trainable}
    gnorm = clip_and_sgd(params, grads, mom, lr, clip=clip)
    for k in trainable:
        state[k] = params[k]
    state.update(new_target)
    state.update(bufs)
    return dict(loss_byol=float(loss_byol), ce=[float(c) for c in ce], loss_total=float(tot), grad_norm=gnorm,
                logits=[l.detach() for l in logits], grads=grads)


# ----------------------------------------------------------------------------------------------- finetune / test
def _bn_eval(x, P, name):
    return F.batch_norm(x, P[name + ".running_mean"], P[name + ".running_var"], P[name + ".weight"], P[name + ".bias"],
                        False, BN_MOMENTUM, BN_EPS)


def finetune_forward(state: dict, x, training: bool = True, tape: Tape | None = None):
    """R21DBYOL(pretrain=False).forward(x, o_type in {'ft_fc','ft_all','test'}) -- r21d_byol.py:394-399:
    classify(cls_bn(F.normalize(online_net(x), p=2, dim=1))).  training=False is model.eval() (running statistics,
    main_ft_mp.py:281-289 / test.py:76-93).  Returns (logits, updated BN buffers)."""
    tape = tape or Tape(False)
    if training:
        bufs: dict = {}
        P = _ParamView(state, bufs)
        feat = r2plus1d_net(x, P, "online_net", bufs, tape, "online.v1", proj=False)
        feat = F.normalize(feat, p=2, dim=1)
        feat = _bn(feat, P, "cls_bn", bufs, tape, "cls_bn")
        return F.linear(feat, state["classify.weight"], state["classify.bias"]), bufs
    # eval mode: the same graph with every BatchNorm replaced by its running-statistics affine map
    P = state

    def stc(x_, pre, kernel, stride, pad):
        x_ = F.conv3d(x_, P[pre + ".spatial_conv.weight"], None, (1, stride[1], stride[2]), (0, pad[1], pad[2]))
        x_ = F.relu(_bn_eval(x_, P, pre + ".bn"))
        return F.conv3d(x_, P[pre + ".temporal_conv.weight"], None, (stride[0], 1, 1), (pad[0], 0, 0))
    h = stc(x, "online_net.conv1", (3, 7, 7), (1, 2, 2), (1, 3, 3))
    h = F.relu(_bn_eval(h, P, "online_net.bn1"))
    for stage, down in (("conv2", False), ("conv3", True), ("conv4", True), ("conv5", True)):
        pre = f"online_net.{stage}.block1"
        s_ = (2, 2, 2) if down else (1, 1, 1)
        r = stc(h, pre + ".conv1", (3, 3, 3), s_, (1, 1, 1))
        r = F.relu(_bn_eval(r, P, pre + ".bn1"))
        r = _bn_eval(stc(r, pre + ".conv2", (3, 3, 3), (1, 1, 1), (1, 1, 1)), P, pre + ".bn2")
        if down:
            h = _bn_eval(stc(h, pre + ".downsampleconv", (1, 1, 1), (2, 2, 2), (0, 0, 0)), P, pre + ".downsamplebn")
        h = F.relu(h + r)
    feat = F.adaptive_avg_pool3d(h, 1).view(-1, 512)
    feat = _bn_eval(F.normalize(feat, p=2, dim=1), P, "cls_bn")
    return F.linear(feat, state["classify.weight"], state["classify.bias"]), {}


def finetune_step(state: dict, trainable: list[str], x, labels, lr, mom: dict, momentum=0.9, wd=1e-3, tape=None):
    """One step of main_ft_mp.py:196-214 (CrossEntropyLoss, backward, SGD.step; no gradient clipping)."""
    for k in trainable:
        state[k] = state[k].detach().requires_grad_(True)
    logits, bufs = finetune_forward(state, x, True, tape)
    loss = F.cross_entropy(logits, labels)
    grads = dict(zip(trainable, torch.autograd.grad(loss, [state[k] for k in trainable], allow_unused=True)))
    params = {k: state[k].detach() for k in trainable}
    clip_and_sgd(params, grads, mom, lr, momentum=momentum, wd=wd, clip=False)
    for k in trainable:
        state[k] = params[k]
    state.update(bufs)
    return dict(loss=float(loss), logits=logits.detach(), grads=grads)


# ----------------------------------------------------------------------------------------------- NT-Xent
def ntxent_reference_form(zis, zjs, temperature, use_cosine=True):
    """loss/NTXent.py:46-62 restated literally (materialises the 2N x 2N matrix; small N only)."""
    n = zis.shape[0]
    reps = torch.cat([zjs, zis], dim=0)
    if use_cosine:
        sim = F.cosine_similarity(reps.unsqueeze(1), reps.unsqueeze(0), dim=-1)
    else:
        sim = reps @ reps.t()
    l_pos = torch.diag(sim, n)
    r_pos = torch.diag(sim, -n)
    positives = torch.cat([l_pos, r_pos]).view(2 * n, 1)
    mask = ~(torch.eye(2 * n, dtype=torch.bool) | torch.eye(2 * n, dtype=torch.bool).roll(n, 1))
    negatives = sim[mask].view(2 * n, -1)
    logits = torch.cat((positives, negatives), dim=1) / temperature
    labels = torch.zeros(2 * n, dtype=torch.long)
    return F.cross_entropy(logits, labels, reduction="sum") / (2 * n)


def ntxent_closed_form(zis, zjs, temperature, use_cosine=True):
    """mean_i[LSE_{j != i}(s_ij/tau) - s_i,pos(i)/tau], O(rows^2) memory, no rows^2*d broadcast."""
    n = zis.shape[0]
    z = torch.cat([zjs, zis], dim=0)
    if use_cosine:
        z = z / z.norm(dim=1, keepdim=True).clamp_min(1e-8)
    s = z @ z.t() / temperature
    idx = torch.arange(2 * n)
    pos = (idx + n) % (2 * n)
    sm = s.masked_fill(torch.eye(2 * n, dtype=torch.bool), float("-inf"))
    return (torch.logsumexp(sm, dim=1) - s[idx, pos]).mean()


# ----------------------------------------------------------------------------------------------- protocol
# The seeded input protocol (no algorithm in it) lives in cstp_b200/synthetic.py so that bench.py's native arm and the
# profiling tools can draw the same inputs without importing the oracle; re-exported here for the tests.
from cstp_b200.synthetic import structured_batch, synthetic_batch  # noqa: E402,F401

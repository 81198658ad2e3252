"""`generate_model(opts)` with the contract of the reference's models/model.py:39-144 for the `r21d_byol` branch.

Only the pretraining tasks of the north-star hot path are wired to the B200 engine; every other backbone / task of
the reference factory is outside the scope table and raises.
"""
from __future__ import annotations

import torch
from torch import nn

from .pace import r21d_byol


def generate_model(opts):
    if opts.model_name != "r21d_byol":
        raise ValueError("Please check the input backbone!")          # models/model.py:79
    if opts.task not in ("r_byol", "loss_com", "resume"):
        raise NotImplementedError(f"task {opts.task!r} is outside the pretraining hot path implemented by cstp_b200")
    model = r21d_byol.R21DBYOL(pretrain=True)
    if getattr(opts, "distributed", False):
        # models/model.py:82-103.  The reference's --sync_bn builds a process group holding only the local rank, i.e.
        # per-GPU statistics; the engine's BatchNorm is per-GPU as well, so both flags map to plain DDP.
        torch.cuda.set_device(opts.local_rank)
        model.cuda(opts.local_rank)
        model = nn.parallel.DistributedDataParallel(model, device_ids=[opts.local_rank], output_device=opts.local_rank,
                                                    find_unused_parameters=False, broadcast_buffers=False)
    else:
        model = model.to(opts.device)
    if opts.task == "resume":
        md = torch.load(opts.resume_md_path, map_location="cpu")
        assert opts.arch == md["arch"]
        state = {k[len("module."):] if k.startswith("module.") and not isinstance(model, nn.parallel.DistributedDataParallel)
                 else k: v for k, v in md["state_dict"].items()}
        model.load_state_dict(state)
    return model, model.parameters()

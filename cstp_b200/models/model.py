"""`generate_model(opts)` with the contract of the reference's models/model.py:39-144 for the `r21d_byol` branch:
pretraining (`r_byol`, `loss_com`, `resume`), finetuning (`ft_fc`, `ft_all`, `scratch`) and `test`, including the
pretrain -> finetune checkpoint hand-off of `neq_load_customized` (models/model.py:11-36).  Every other backbone of the
reference factory is outside the scope table and raises the reference's own ValueError.
"""
from __future__ import annotations

import torch
from torch import nn

from .pace import r21d_byol


def neq_load_customized(model, pretrained_dict, verbose=True):
    """models/model.py:11-36: copy every pretrained entry whose key also exists in the new model (the `online_net.*`
    backbone when going from the pretraining to the finetune model), keep the rest of the new model as initialised."""
    model_dict = model.state_dict()
    tmp = {k: v for k, v in pretrained_dict.items() if k in model_dict}
    if verbose:
        print("\n=======Check Weights Loading======")
        print("Weights not loaded into new model:")
        for k in model_dict:
            if k not in pretrained_dict:
                print(k)
        print("===================================\n")
    model_dict.update(tmp)
    model.load_state_dict(model_dict)
    return model


def generate_model(opts):
    if opts.model_name != "r21d_byol":
        raise ValueError("Please check the input backbone!")          # models/model.py:79
    if opts.task in ("r_byol", "loss_com", "resume"):
        model = r21d_byol.R21DBYOL(pretrain=True)
    elif opts.task in ("ft_fc", "ft_all", "scratch", "test"):
        model = r21d_byol.R21DBYOL(pretrain=False, num_classes=opts.n_classes, cls_bn=True)      # models/model.py:48-49
    else:
        raise NotImplementedError(f"task {opts.task!r} is outside the r21d_byol paths implemented by cstp_b200")
    if getattr(opts, "bn_sync", "local") == "world":
        # opt-in extension (not a reference flag): BatchNorm statistics over every rank of the default process group
        from ..parallel import BnSync
        model.engine_options = {"bn_sync": BnSync()}
    wrapped = False
    if getattr(opts, "distributed", False):
        # models/model.py:82-103.  The reference's --sync_bn builds a process group holding only the local rank, i.e.
        # per-GPU statistics; the engine's BatchNorm is per-GPU as well, so both flags map to plain DDP.
        # DDP defaults as in the reference (broadcast_buffers stays True: rank 0's BatchNorm running statistics are what
        # every rank -- and the checkpoint -- carries).
        if torch.cuda.is_available():
            torch.cuda.set_device(opts.local_rank)
            model.cuda(opts.local_rank)
            model = nn.parallel.DistributedDataParallel(model, device_ids=[opts.local_rank], output_device=opts.local_rank,
                                                        find_unused_parameters=(opts.task == "ft_fc"))
        else:
            # no GPU in this process: only the host logic can run (tests drive it over a CPU stand-in of the kernel layer;
            # the engine itself refuses CPU tensors)
            model = nn.parallel.DistributedDataParallel(model, find_unused_parameters=(opts.task == "ft_fc"))
        wrapped = True
    else:
        model = model.to(opts.device)

    def strip(sd):
        """Checkpoints are saved from the DDP-wrapped model (`module.` prefix, main_byol.py:137)."""
        return {(k[len("module."):] if k.startswith("module.") and not wrapped else k): v for k, v in sd.items()}

    if opts.task in ("scratch", "r_byol", "loss_com"):
        return model, model.parameters()
    if "test" in opts.task:
        md = torch.load(opts.test_md_path, map_location=opts.device)
        assert opts.arch == md["arch"]
        model.load_state_dict(strip(md["state_dict"]))
        return model
    if opts.task == "resume":
        md = torch.load(opts.resume_md_path, map_location="cpu")
        assert opts.arch == md["arch"]
        model.load_state_dict(strip(md["state_dict"]))
        return model, model.parameters()
    # ft_fc / ft_all: models/model.py:122-142
    opts.ft_begin_index = 5 if opts.task == "ft_fc" else 0
    ck = torch.load(opts.pretrained_path, map_location=torch.device("cpu"))
    assert opts.arch in ck["arch"] or ck["arch"] in opts.arch
    model = neq_load_customized(model, strip(ck["state_dict"]), verbose=False)
    return model, r21d_byol.get_fine_tuning_parameters(model, opts.ft_begin_index)

"""CPU: host-side logic of the training path — packed-gradient buffer layout, the un-packing into the
reference's state_dict layout (qkv row interleave, subsampling.out column order), and the data-parallel
gradient all-reduce over a world_size-2 gloo group (the N>1 path of cfg 5 without a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lcasr_oracle as O

CFG = dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127)


def _engine():
    import lcasr_b200
    from lcasr_b200.training import TrainEngine
    model = lcasr_b200.SCConformerXL(**O.make_config(**CFG))
    model.load_state_dict(O.synth_state_dict(O.make_config(**CFG), seed=5), strict=True)
    return model, TrainEngine(model)


def test_packed_gradient_layout_round_trip():
    """a gradient written in PACKED layout comes back in the parameter's own layout: d(sum(w_packed * r)) = r"""
    model, eng = _engine()
    P = eng.pack(torch.device("cpu"))
    flat, G = eng.grad_buffers(P, torch.device("cpu"))
    assert flat.numel() % 4 == 0 and all(v.data_ptr() % 16 == 0 for v in G.values())
    g = torch.Generator().manual_seed(0)
    for k in G:
        G[k].copy_(torch.randn(G[k].shape, generator=g))
    grads = eng.to_param_grads(G)
    params = dict(model.named_parameters())
    assert set(grads) == set(params)
    # reference: autograd through the same packing applied to the fp32 parameters
    H, Dh, d, Cc, F3 = model.n_heads, model.head_dim, model.d_model, model.subsampling_conv_channels, model.subsampling.feat_sub
    for k, key in eng._names.items():
        p = params[key].detach().clone().requires_grad_(True)
        if k == "sub_out_w":
            packed = p.reshape(d, Cc, F3).permute(0, 2, 1).reshape(d, F3 * Cc)
        elif k.endswith("qkv_w"):
            packed = p.reshape(H, Dh, 3, d).permute(2, 0, 1, 3).reshape(3 * H * Dh, d)
        else:
            packed = p.reshape(G[k].shape)
        (packed * G[k]).sum().backward()
        assert torch.equal(p.grad, grads[key]), key


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, eng = _engine()
        eng.dp_group = dist.group.WORLD
        P = eng.pack(torch.device("cpu"))
        flat, G = eng.grad_buffers(P, torch.device("cpu"))
        for i, k in enumerate(G):
            G[k].fill_(float(rank + 1) * (i + 1))
        eng.reduce_gradients(flat)
        ok = all(torch.allclose(G[k], torch.full_like(G[k], (i + 1) * (1 + world) / 2.0)) for i, k in enumerate(G))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_train_mode_rejects_what_it_does_not_cover():
    import lcasr_b200
    model = lcasr_b200.SCConformerXL(**O.make_config(**CFG)).train()
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        model(torch.zeros(1, 80, 64))

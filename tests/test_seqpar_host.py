"""CPU: host logic of the sequence-parallel path — token/frame partitioning and the torch.distributed
plumbing (world_size 2, gloo backend, 127.0.0.1), no kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lcasr_oracle as O


def test_split_and_frame_slices_cover_the_recording():
    from lcasr_b200 import seqpar
    for T, world in [(131072, 8), (360000, 8), (1024, 2), (2056, 3), (360000, 7)]:
        N = O.calc_length(T)
        blocks = seqpar.split_tokens(N, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(e - s for s, e in blocks) - min(e - s for s, e in blocks) <= 1
        for r, (s, e) in enumerate(blocks):
            f0, f1, drop = seqpar.frame_slice(s, e)
            assert drop == (1 if r > 0 else 0)
            assert f0 == 8 * (s - drop) and f1 == 8 * e and 0 <= f0 < f1 <= T
            # the slice subsamples to exactly the owned tokens plus the dropped context token
            assert O.calc_length(f1 - f0) == (e - s) + drop


def test_subsampling_slice_equals_full_subsampling_on_cpu():
    """The no-communication claim, checked with the oracle: subsampling a slice with one token of left
    context and dropping that token reproduces the full-recording subsampling exactly."""
    from lcasr_b200 import seqpar
    cfg = O.make_config(n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127)
    sd = O.synth_state_dict(cfg, seed=3)
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    x = O.synth_input(1, 1024, seed=4)
    full = O.subsampling_forward(sd, cfg, x)
    for world in (2, 3, 4):
        parts = []
        for s, e in seqpar.split_tokens(full.shape[1], world):
            f0, f1, drop = seqpar.frame_slice(s, e)
            parts.append(O.subsampling_forward(sd, cfg, x[:, :, f0:f1])[:, drop:])
        assert (torch.cat(parts, 1) - full).abs().max().item() < 1e-5


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ragged):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lcasr_b200 import seqpar
    comm = seqpar.DistComm()
    n = 5 + (rank if ragged else 0)
    x = torch.arange(n * 3, dtype=torch.float32).view(n, 3) + 100 * rank
    (g,) = comm.all_gather_cat([x])
    exp = torch.cat([torch.arange((5 + (r if ragged else 0)) * 3, dtype=torch.float32).view(-1, 3) + 100 * r for r in range(world)])
    assert torch.equal(g, exp)
    first, last = x[:2].clone(), x[-2:].clone()
    ((left, right),) = comm.halo_exchange([first], [last])
    if rank == 0:
        assert torch.equal(left, torch.zeros_like(first))
    else:
        prev = torch.arange((5 + ((rank - 1) if ragged else 0)) * 3, dtype=torch.float32).view(-1, 3) + 100 * (rank - 1)
        assert torch.equal(left, prev[-2:])
    if rank == world - 1:
        assert torch.equal(right, torch.zeros_like(last))
    else:
        nxt = torch.arange((5 + ((rank + 1) if ragged else 0)) * 3, dtype=torch.float32).view(-1, 3) + 100 * (rank + 1)
        assert torch.equal(right, nxt[:2])
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_dist_comm_gloo_world2(ragged):
    mp.spawn(_worker, args=(2, _free_port(), ragged), nprocs=2, join=True)


def test_local_comm_matches_dist_semantics():
    from lcasr_b200 import seqpar
    comm = seqpar.LocalComm(3)
    xs = [torch.full((2, 2), float(r)) for r in range(3)]
    g = comm.all_gather_cat(xs)
    assert len(g) == 3 and torch.equal(g[1], torch.cat(xs))
    halos = comm.halo_exchange([x[:1] for x in xs], [x[-1:] for x in xs])
    assert torch.equal(halos[0][0], torch.zeros(1, 2)) and torch.equal(halos[0][1], xs[1][:1])
    assert torch.equal(halos[2][0], xs[1][-1:]) and torch.equal(halos[2][1], torch.zeros(1, 2))

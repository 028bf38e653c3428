"""world_size-2 gloo test (CPU) of the data-parallel plumbing: clip sharding, the flat-gradient
all-reduce and the 1/world scaling give the same update as one process averaging both clips'
gradients.  Gradients come from the oracle (small T) -- the device kernels are not involved."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(d, keys):
    return torch.cat([d[k].reshape(-1) for k in keys])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import encoder_oracle as eo
    from zeronotesamba_b200 import dist_utils
    torch.set_num_threads(2)
    assert dist_utils.is_distributed()
    clips = dist_utils.shard_clips(5, rank, world)
    assert clips == [rank, rank + world]
    sd = eo.he_normal_state_dict(3)
    keys = list(sd.keys())
    flat_p = _flat(sd, keys).clone()
    if rank == 1:
        flat_p += 1.0                      # replicas must start from rank 0's weights
    dist_utils.broadcast_parameters(flat_p, 0)
    assert torch.equal(flat_p, _flat(sd, keys))
    g = torch.Generator().manual_seed(50 + clips[0])
    batch = (torch.rand(2, 2, 96, 12, generator=g) * 10 - 9)
    res = eo.pretext_step(sd, batch, batch_len=2, temperature=0.25)
    flat_g = _flat(res["grads"], keys).clone()
    scale = dist_utils.allreduce_gradients(flat_g)
    assert scale == 0.5
    m = np.zeros(flat_p.numel()); v = np.zeros(flat_p.numel())
    new_p, _, _ = eo.adam_step(flat_p.numpy().astype(np.float64), (flat_g * scale).numpy().astype(np.float64), m, v, 1)
    np.save(os.path.join(out_dir, f"p{rank}.npy"), new_p)
    np.save(os.path.join(out_dir, f"g{rank}.npy"), _flat(res["grads"], keys).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    assert np.array_equal(p0, p1), "replicas diverged"
    sys.path.insert(0, ROOT)
    from oracle import encoder_oracle as eo
    sd = eo.he_normal_state_dict(3)
    keys = list(sd.keys())
    g_mean = 0.5 * (np.load(tmp_path / "g0.npy").astype(np.float64) + np.load(tmp_path / "g1.npy").astype(np.float64))
    flat_p = _flat(sd, keys).numpy().astype(np.float64)
    want, _, _ = eo.adam_step(flat_p, g_mean, np.zeros_like(flat_p), np.zeros_like(flat_p), 1)
    assert np.allclose(p0, want, rtol=0, atol=1e-9)


def test_shard_clips_partition():
    from zeronotesamba_b200.dist_utils import shard_clips
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_clips(28800, r, world))
        assert seen == list(range(28800 // world * world))
        assert len({len(shard_clips(28801, r, world)) for r in range(world)}) == 1

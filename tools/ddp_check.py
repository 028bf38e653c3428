"""torchrun --nproc-per-node N tools/ddp_check.py : data-parallel sanity on real GPUs.
Each rank trains on its own synthetic source clip (VQT in the loop, dropout off); after 3 steps the flat
parameter buffers must be bit-identical across ranks and match a single-process emulation that sums
the ranks' gradients by hand (every rank runs the emulation: no collective is issued inside it)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from zeronotesamba_b200 import _lib as L
from zeronotesamba_b200 import synth
from zeronotesamba_b200.models.checkpoint import he_normal_state_dict
from zeronotesamba_b200.models.models import Pretext_CNN
from zeronotesamba_b200.pretext import PretextTrainer


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    B = 16

    def inputs(r, step):
        drums, other = synth.stem_pair(100 * r + step, 10.0)
        starts = random.Random(100 * r + step).sample(range(313), B)
        return (torch.from_numpy(other).to(dev), torch.from_numpy(drums).to(dev),
                torch.tensor(starts, dtype=torch.int32, device=dev))

    model = Pretext_CNN().to(dev)
    model.load_state_dict(he_normal_state_dict(7))
    tr = PretextTrainer(model, batch_len=B, dropout_p=0.0, use_graph=True, lr=1e-4)
    for step in range(3):
        res = tr.step_from_audio(*inputs(rank, step))
    torch.cuda.synchronize()
    # the NCCL all-reduce + local Adam path must agree with the fused peer-memory optimizer
    m3 = Pretext_CNN().to(dev)
    m3.load_state_dict(he_normal_state_dict(7))
    tr3 = PretextTrainer(m3, batch_len=B, dropout_p=0.0, use_graph=True, lr=1e-4, p2p_adam=False)
    for step in range(3):
        tr3.step_from_audio(*inputs(rank, step))
    torch.cuda.synchronize()
    d_nccl = (tr3.flat_p - tr.flat_p).abs().max().item()
    gathered = [torch.empty_like(tr.flat_p) for _ in range(world)]
    dist.all_gather(gathered, tr.flat_p)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    # emulation (no collectives): distributed=False trainer, gradients of all ranks summed by hand
    m2 = Pretext_CNN().to(dev)
    m2.load_state_dict(he_normal_state_dict(7))
    tr2 = PretextTrainer(m2, batch_len=B, dropout_p=0.0, use_graph=False, lr=1e-4, distributed=False)
    for step in range(3):
        acc = torch.zeros_like(tr2.flat_g)
        for r in range(world):
            a, p, st = inputs(r, step)
            tr2.step_from_audio(a, p, st, run_step=False)
            if r > 0:
                tr2.engine.step_ctr -= 1       # same step number for every emulated rank
            tr2._forward_backward()
            acc += tr2.flat_g
        tr2.flat_g.copy_(acc)
        L.check(L.lib().zns_adam_flat(L.ptr(tr2.flat_p), L.ptr(tr2.flat_g), L.ptr(tr2.flat_m), L.ptr(tr2.flat_v),
                                      tr2.flat_p.numel(), tr2.lr, 0.9, 0.999, 1e-8, 0, L.ptr(tr2.engine.step_ctr),
                                      1.0 / world, L.current_stream()))
    torch.cuda.synchronize()
    d = (tr2.flat_p - tr.flat_p).abs().max().item()
    p0 = torch.cat([v.reshape(-1) for v in he_normal_state_dict(7).values()]).to(dev)
    moved = (tr.flat_p[:p0.numel()] - p0).abs().max().item()      # (the flat buffer is padded to a multiple of the world size)
    # fp32 atomics / NCCL sum in another order: where a gradient is ~0 Adam's first steps flip sign, 2*lr per step
    ok = same and d < 3 * 2e-4 * 1.05 and d_nccl < 1e-3 and moved > 1e-5 and res[0].item() == res[0].item()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ddp_check world={world}: optimizer={'fused peer-memory (zns_adam_p2p)' if tr._symm is not None else 'NCCL all-reduce + local Adam'} "
              f"max|p_p2p - p_nccl|={d_nccl:.3e}")
        print(f"ddp_check world={world}: replicas identical={same} max|p_ddp - p_emulated|={d:.3e} (3 Adam steps of 1e-4; "
              f"fp32 atomics order differs) moved={moved:.3e} loss/cos={res.tolist()}")
        print("DDP_CHECK_OK" if int(flag.item()) == 1 else "DDP_CHECK_FAILED")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

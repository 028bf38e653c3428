"""torchrun --nproc-per-node N tools/ddp_check.py : data-parallel correctness on real GPUs.

Each rank trains on its own synthetic source clip (VQT in the loop, dropout off) at a well-conditioned operating point
(tied branch weights, anchor stem = drums + 5 % other: cos+ - cos- = 0.04 .. 0.09 depending on the clip, so the gradients are far from the softmax's
flat spot).  Checks, all of which a missing or wrong gradient exchange fails:

  1. exchange: after step 1 Adam's first moment is (1 - beta1) * mean over ranks of the rank gradients.  Every rank
     recomputes every rank's step-1 gradient locally (no collective) and compares its moment buffer with that mean to
     1e-5 of the buffer's largest entry (the fused kernel keeps moments only for the shard it owns: the comparison runs
     over the owned shard; the NCCL path keeps all of them: whole buffer).
  2. replicas: after 3 steps the flat parameter buffers are bit-identical on all ranks.
  3. both optimizers: fused peer-memory kernel (zns_adam_p2p) and NCCL all-reduce + local Adam agree after 3 steps to a few
     sign flips of near-zero gradient entries (3 steps x 2 lr), and the parameters really moved (> lr on average).
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from zeronotesamba_b200 import synth
from zeronotesamba_b200.models.checkpoint import he_normal_state_dict
from zeronotesamba_b200.models.models import Pretext_CNN
from zeronotesamba_b200.pretext import PretextTrainer

LR, BETA1, B, STEPS = 1e-4, 0.9, 16, 3


def tied(sd):
    out = {k: v.clone() for k, v in sd.items()}
    for k in list(out):
        if k.startswith("postve."):
            out[k] = out["anchor." + k[len("postve."):]].clone()
    return out


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    sd = tied(he_normal_state_dict(7))

    def inputs(r, step):
        drums, other = synth.stem_pair(100 * r + step, 10.0)
        anchor = np.clip(drums + 0.05 * other, -0.9, 0.9).astype(np.float32)
        starts = random.Random(100 * r + step).sample(range(313), B)
        return (torch.from_numpy(anchor).to(dev), torch.from_numpy(drums).to(dev),
                torch.tensor(starts, dtype=torch.int32, device=dev))

    def trainer(**kw):
        m = Pretext_CNN().to(dev)
        m.load_state_dict(sd)
        return PretextTrainer(m, batch_len=B, dropout_p=0.0, lr=LR, **kw)

    # ---- every rank's step-1 gradient, computed locally without any collective ---------------------------------------
    solo = trainer(use_graph=False, distributed=False)
    n = solo.flat_g.numel()
    g_mean = torch.zeros(n, device=dev, dtype=torch.float64)
    for r in range(world):
        a, p, st = inputs(r, 0)
        solo.step_from_audio(a, p, st, run_step=False)
        solo._forward_backward()
        g_mean += solo.flat_g.double()
    g_mean /= world
    torch.cuda.synchronize()
    res_solo = solo.result.tolist()
    p0 = solo.flat_p.clone()                     # same layout (padding included) as the distributed trainers' buffers

    report = {}
    finals = {}
    for name, kw in (("fused", dict(p2p_adam=True)), ("nccl", dict(p2p_adam=False))):
        tr = trainer(use_graph=True, **kw)
        assert tr.flat_p.numel() >= n
        path = ("fused peer-memory reduce-scatter + Adam + all-gather (zns_adam_p2p)" +
                (", late bucket exchanged under the early half of the backward pass" if tr.dp_overlap else ", after the backward pass")) \
            if tr._symm is not None else "NCCL all-reduce + local Adam (zns_adam_flat)"
        tr.step_from_audio(*inputs(rank, 0))
        torch.cuda.synchronize()
        m = tr.flat_m[:n].double()
        want = (1.0 - BETA1) * g_mean
        if tr._symm is not None:                 # moments exist for the owned shards only
            # zns_adam_p2p splits the float4 index space of every range it is called on evenly over the ranks; the overlapped
            # step calls it per bucket range (late / early, per branch), the plain step once for the whole buffer
            ranges = (tr._bucket_late + tr._bucket_early) if tr.dp_overlap else [(0, tr.flat_p.numel())]
            owned = []
            for r_lo, r_hi in ranges:
                n4 = (r_hi - r_lo) // 4
                per = (n4 + world - 1) // world
                owned.append((r_lo + min(4 * per * rank, 4 * n4), r_lo + min(4 * per * (rank + 1), 4 * n4)))
        else:
            owned = [(0, n)]
        sel = torch.zeros(n, dtype=torch.bool, device=dev)
        for o_lo, o_hi in owned:
            sel[o_lo:min(o_hi, n)] = True
        lo, hi = int(sel.sum()), len(owned)      # reported: elements checked, number of shards
        scale = float(want.abs().max())
        err_m = float((m[sel] - want[sel]).abs().max()) / scale
        # the moment must be the MEAN: against a single rank's gradient the same comparison is far off
        own = (1.0 - BETA1) * solo_grad_of(solo, inputs, rank)
        err_own = float((m[sel] - own[:n][sel].double()).abs().max()) / scale if world > 1 else float("nan")
        for step in range(1, STEPS):
            tr.step_from_audio(*inputs(rank, step))
        torch.cuda.synchronize()
        gathered = [torch.empty_like(tr.flat_p) for _ in range(world)]
        dist.all_gather(gathered, tr.flat_p.contiguous())
        same = all(torch.equal(gathered[0], g) for g in gathered)
        moved = float((tr.flat_p[:n] - p0).abs().mean())
        finals[name] = tr.flat_p[:n].clone()
        report[name] = dict(path=path, err_m=err_m, err_own=err_own, same=same, moved=moved, res=tr.result.tolist(),
                            shard=(lo, hi))
        del tr
    d_paths = float((finals["fused"] - finals["nccl"]).abs().max())
    ok = True
    for name, r in report.items():
        ok &= r["err_m"] < 1e-5 and r["same"] and r["moved"] > 0.5 * LR
        if world > 1:
            ok &= r["err_own"] > 1e-3            # the check above is discriminating: one rank's gradient does not pass it
    ok &= d_paths < STEPS * 2 * LR * 1.05
    ok &= res_solo[1] - res_solo[2] > 0.03        # non-degenerate operating point (uniform softmax would be cos+ == cos-)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ddp_check world={world}: operating point loss/cos+/cos- = {res_solo}")
        for name, r in report.items():
            print(f"ddp_check world={world} [{name}]: exchange path = {r['path']}")
            print(f"    step-1 first moment vs (1-beta1)*mean(rank gradients), rank 0: {r['shard'][0]} elements in {r['shard'][1]} owned shard(s): max err / max|m| = {r['err_m']:.3e}"
                  f"  (vs rank 0's own gradient: {r['err_own']:.3e});  replicas bit-identical after {STEPS} steps = {r['same']};"
                  f"  mean |p - p0| = {r['moved']:.3e} (lr {LR:g})")
        print(f"ddp_check world={world}: max |p_fused - p_nccl| after {STEPS} steps = {d_paths:.3e} (bound {STEPS * 2 * LR * 1.05:.2e}: "
              f"sign flips of near-zero gradient entries under a different summation order)")
        print("DDP_CHECK_OK" if int(flag.item()) == 1 else "DDP_CHECK_FAILED")
    dist.barrier()
    dist.destroy_process_group()


def solo_grad_of(solo, inputs, r):
    a, p, st = inputs(r, 0)
    solo.step_from_audio(a, p, st, run_step=False)
    solo._forward_backward()
    torch.cuda.synchronize()
    return solo.flat_g.clone()


if __name__ == "__main__":
    main()

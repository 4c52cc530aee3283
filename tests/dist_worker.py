"""Worker of tests/test_dist_gpu.py (one process per GPU under torch.distributed.run): data-parallel DCCRN train steps on real
GPUs over NCCL.  Checks, on every rank: (1) ranks that were constructed from DIFFERENT seeds hold bit-identical parameters after
construction of TrainStep (rank-0 broadcast) and after 3 steps on different shards; (2) the all-reduced gradient times the Adam
scale equals the mean of the per-rank CUDA gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")]

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import models
    from sefd import dist as sdist
    from sefd.train import TrainStep
    models.cfg.loss = "SI-SNR"
    torch.manual_seed(100 + rank)                       # every rank starts from a different random init
    m = models.DCCRN(masking_mode="C").to(dev).train()
    ts = TrainStep(m, lr=1e-3, loss="SI-SNR")
    eng = ts.engine

    def same_everywhere(t, what):
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, t), f"rank {rank}: {what} differs from rank 0"

    same_everywhere(eng.flat, "parameters after TrainStep construction")
    g = torch.Generator().manual_seed(7)
    B, L = 4, 8000
    noisy = ((torch.rand(world * B, L, generator=g) * 2 - 1) * 0.1)
    clean = ((torch.rand(world * B, L, generator=g) * 2 - 1) * 0.1)
    lo, hi = sdist.shard_rows(world * B, rank, world)
    xn, xc = noisy[lo:hi].contiguous().to(dev), clean[lo:hi].contiguous().to(dev)
    # the reduced gradient is the mean of the shard gradients
    ts.forward_backward(xn, xc)
    local_grad = eng.flat_grad.clone()
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    mean = torch.stack(gathered).double().mean(0)
    scale = sdist.allreduce_sum_(eng.flat_grad)
    got = eng.flat_grad.double() * scale
    err = float((got - mean).abs().max() / (mean.abs().max() + 1e-30))
    assert err < 1e-6, f"rank {rank}: reduced gradient vs mean of shard gradients: rel err {err}"
    # the overlapped reduction (tail slice all-reduced on a side stream beside the encoder backward, head slice after it) yields
    # the same gradient as one all-reduce after the backward (compared on the gradient itself: Adam turns rounding-level noise
    # of near-zero gradients into +-lr updates, so parameters after a few steps are not a usable comparison)
    assert ts.overlap, "the tail-slice all-reduce overlap should be on for a multi-rank DCCRN step"
    ts.forward_backward(xn, xc, reduce_tail=True)
    assert ts._tail_pending
    sdist.allreduce_sum_(eng.flat_grad[:ts._split])
    torch.cuda.current_stream().wait_stream(ts._comm_stream)
    ts._tail_pending = False
    got2 = eng.flat_grad.double() * scale
    err2 = float((got2 - mean).abs().max() / (mean.abs().max() + 1e-30))
    assert err2 < 1e-6, f"rank {rank}: overlapped reduction vs mean of shard gradients: rel err {err2}"
    for _ in range(3):
        ts.step(xn, xc)
    same_everywhere(eng.flat, "parameters after 3 data-parallel steps")
    same_everywhere(ts.exp_avg, "Adam first moment after 3 steps")
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world} grad_rel_err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

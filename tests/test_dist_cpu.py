"""World-size-2 gloo test (CPU) of the multi-rank logic: batch sharding + one sum all-reduce of the flat gradient
buffer + 1/world scaling reproduce the gradient of the mean loss over the global batch with per-rank BatchNorm
statistics (DDP semantics).  The per-rank gradients come from the CPU oracle."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dccrn_oracle as O
    from sefd import dist as D
    torch.set_num_threads(2)
    sd0 = O.init_state(0)
    noisy, clean = O.synthetic_batch(4, 1000)
    lo, hi = D.shard_rows(4)
    tr = O.OracleTrainer(sd0)
    tr.forward_backward(noisy[lo:hi], clean[lo:hi])
    keys = tr.keys
    flat = torch.cat([tr.sd[k].grad.reshape(-1) for k in keys])
    local = flat.clone()
    scale = D.allreduce_sum_(flat)
    torch.save({"rank": rank, "lo": lo, "hi": hi, "local": local, "reduced": flat * scale, "scale": scale},
               os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 2, 2, 4)
    assert r0["scale"] == 0.5
    assert torch.equal(r0["reduced"], r1["reduced"])
    torch.testing.assert_close(r0["reduced"], 0.5 * (r0["local"] + r1["local"]), rtol=1e-6, atol=1e-9)
    assert float((r0["local"] - r1["local"]).abs().max()) > 0          # the shards really differ


def test_shard_rows_covers_everything():
    from sefd import dist as D
    for n in (1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            cuts = [D.shard_rows(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1

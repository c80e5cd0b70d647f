"""torchrun worker for tests/test_gpu_sharded.py: sharded search over NCCL must
equal the single-GPU search bit for bit (the merge is pure selection)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloudvectordb_b200 import IndexFlat, IndexIVFFlat, Kmeans, ShardedIndex, ShardedIVFFlat, mine_hard_negatives, mine_hard_negatives_sharded  # noqa: E402
from cloudvectordb_b200.sharded import shard_bounds  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device="cpu").manual_seed(99)
    n, d, nq, k = 300_001, 256, 777, 10
    xb = torch.nn.functional.normalize(torch.randn((n, d), generator=g), dim=1).bfloat16()
    xq = torch.nn.functional.normalize(torch.randn((nq, d), generator=g), dim=1).bfloat16()
    groups = (torch.arange(n) // 4).to(torch.int32)
    self_ids = torch.randint(0, n, (nq,), generator=g)
    for metric in ("ip", "l2"):
        full = IndexFlat(d, metric, "bf16", device=local)
        full.add(xb.to(dev))
        full.set_groups(groups)
        D_ref, I_ref = full.search(xq.to(dev), k)
        D2_ref, I2_ref = full.search(xq.to(dev), k, self_ids=self_ids, group_q=groups[self_ids])
        full.close()
        sh = ShardedIndex(d, metric, "bf16", device=local)
        sh.add(xb.to(dev))
        lo, hi = shard_bounds(n, world, rank)
        assert sh.local.ntotal == hi - lo and sh.ntotal == n
        sh.set_groups_local(groups[lo:hi])
        D, I = sh.search(xq.to(dev), k)
        assert torch.equal(I, I_ref) and torch.equal(D, D_ref), f"{metric}: sharded != single GPU"
        D2, I2 = sh.search(xq.to(dev), k, self_ids=self_ids, group_q=groups[self_ids].to(dev))
        assert torch.equal(I2, I2_ref) and torch.equal(D2, D2_ref), f"{metric}: sharded exclusion != single GPU"
        Dh, Ih = sh.search(xq, k)                          # host queries -> host results
        assert not Dh.is_cuda and torch.equal(Ih, I_ref.cpu())
    # sharded self-join == single-GPU self-join on the owner's rows
    m, dm, km_ = 20_000, 64, 20
    emb = torch.nn.functional.normalize(torch.randn((m, dm), generator=g), dim=1).bfloat16()
    grp = (torch.arange(m) // 4).to(torch.int32)
    D1, I1 = mine_hard_negatives(emb.to(dev), km_, grp.to(dev), device=local, chunk=4096)
    shm = ShardedIndex(dm, "ip", "bf16", device=local)
    lo, hi = shard_bounds(m, world, rank)
    shm.add_local(emb[lo:hi].to(dev))
    shm.set_groups_local(grp[lo:hi])
    D2, I2 = mine_hard_negatives_sharded(shm, emb[lo:hi].to(dev), km_, grp[lo:hi].to(dev), chunk=3000)
    assert torch.equal(I2, I1[lo:hi]) and torch.equal(D2, D1[lo:hi]), "sharded mining != single GPU"
    # sharded SYMMETRIC self-join (every pair of rows scored once in the whole job) == the plain join, on the owner's rows
    from cloudvectordb_b200 import mine_hard_negatives_sharded_symmetric
    st = {}
    D3, I3 = mine_hard_negatives_sharded_symmetric(shm, emb[lo:hi].to(dev), km_, grp[lo:hi].to(dev), chunk=2048, first_chunk=512,
                                                   stats=st)
    assert torch.equal(I3, I1[lo:hi]) and torch.allclose(D3, D1[lo:hi], atol=1e-6), f"sharded symmetric join != plain join {st}"
    # uneven shards, no groups, a chunk size that leaves ragged tails
    shu = ShardedIndex(dm, "ip", "bf16", device=local)
    cut = [0] + [int(m * (r + 1) / world) - (37 * (r + 1) if r + 1 < world else 0) for r in range(world)]
    shu.add_local(emb[cut[rank]:cut[rank + 1]].to(dev))
    D4, I4 = mine_hard_negatives_sharded_symmetric(shu, emb[cut[rank]:cut[rank + 1]].to(dev), km_, None, chunk=1024, first_chunk=256)
    D5, I5 = mine_hard_negatives(emb.to(dev), km_, None, device=local, chunk=4096)
    assert torch.equal(I4, I5[cut[rank]:cut[rank + 1]]), "sharded symmetric join (uneven shards) != plain join"
    # sharded IVF == single-GPU IVF with the same centroids (same lists probed, union of the shards' rows)
    cent = emb[:64].float()
    one = IndexIVFFlat(dm, 64, "ip", device=local)
    one.train(None, centroids=cent)
    one.add(emb.to(dev))
    Dv, Iv = one.search(emb[:500].to(dev), 10, nprobe=5)
    one.close()
    shv = ShardedIVFFlat(dm, 64, "ip", device=local)
    shv.train(None, centroids=cent)
    shv.add_local(emb[lo:hi].to(dev))
    assert shv.ntotal == m
    Dw, Iw = shv.search(emb[:500].to(dev), 10, nprobe=5)
    assert torch.equal(Iw, Iv) and torch.equal(Dw, Dv), "sharded IVF != single GPU IVF"
    shv.close()
    # k-means: sharded points, all-reduced update == single-rank update on all points
    pts = torch.nn.functional.normalize(torch.randn((40_000, 64), generator=g), dim=1).bfloat16()
    cent = pts[:128].float()
    km1 = Kmeans(64, 128, niter=1, device=local)
    km1.centroids = cent.clone().to(dev)
    dist_group = dist.new_group(ranks=[rank])            # single-rank group: no reduction across ranks
    km1.group = dist_group
    km1.step(pts.to(dev))
    lo, hi = shard_bounds(pts.shape[0], world, rank)
    km2 = Kmeans(64, 128, niter=1, device=local)
    km2.centroids = cent.clone().to(dev)
    km2.step(pts[lo:hi].to(dev))
    assert torch.equal(km1.last_counts, km2.last_counts)
    assert torch.allclose(km1.centroids, km2.centroids, rtol=1e-4, atol=1e-6)
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

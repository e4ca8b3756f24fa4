"""Phase timing of symmer_b200.dist.sharded_product under torchrun (bench workload)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from symmer_b200 import ops, dist as sdist
from oracle import pauli_oracle as po

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = ops.device()
dist.init_process_group("nccl", device_id=dev)
a_s, a_c = po.random_operator(1000, 12500, seed=100 + rank)
b_s, b_c = po.random_operator(1000, 10000, seed=7)
a = ops.pack(torch.from_numpy(a_s), 1000); ac = torch.from_numpy(a_c).to(dev)
b = ops.pack(torch.from_numpy(b_s), 1000); bc = torch.from_numpy(b_c).to(dev)
lg = sdist.log2_exact(world)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for rep in range(4):
    dist.barrier(); torch.cuda.synchronize()
    t = [ev()]
    a_full, offsets = sdist.all_gather_rows(a); a_c_full, _ = sdist.all_gather_rows(ac); t.append(ev())
    recs = ops.pair_records(a_full, offsets[rank], offsets[rank + 1], b); t.append(ev())
    part, counts = ops.partition_records(recs, lg); t.append(ev())
    mine = sdist.exchange_records(part, counts); t.append(ev())
    oxz, oc = ops.dedup_records(mine, a_full, a_c_full, b, bc); t.append(ev())
    torch.cuda.synchronize()
    names = ["allgather", "pair", "partition", "exchange", "dedup"]
    ms = [t[i].elapsed_time(t[i + 1]) for i in range(5)]
    if rank == 0 and rep >= 2:
        print(" ".join(f"{n}={m:.2f}" for n, m in zip(names, ms)), "total=%.2f" % sum(ms), flush=True)
    del oxz, oc, mine, part, recs
dist.destroy_process_group()

"""Multi-GPU parity of the sharded mode (one formula, work partitioned over the ranks, NCCL exchange):
torchrun --nproc-per-node N scripts/sharded_parity.py
Every rank runs the same Coprocessor call with CPsetDistributed; scan requests / entry tiles / BVE variable
batches are partitioned, hits, pair sizes and resolvents are all-gathered.  Each rank's ordered result must equal
the CPU oracle's (so it is independent of the partition and of N)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
import riss_solver_b200 as rs
from harness import run_oracle, compare, Result, STAT_NAMES, load_gen

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gen = load_gen()
FULL = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
cases = [("planted 3-SAT 20k", gen.random_ksat(20000, 85200, 3, seed=3, dup=0.03, sup=0.05, flip=0.05)),
         ("community 30k", gen.community_ksat(30000, 150000, seed=20260001))]
bad = 0
for name, (lits, sizes) in cases:
    stream = gen.csr_to_stream(lits, sizes)
    box = [rs.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cp = rs.Coprocessor(FULL + [f"-cp3_gpu_device={local}"])
    cp.add_stream(stream)
    assert cp.set_distributed(rank, world, box[0]) == 0
    dist.barrier(); t0 = time.time()
    cp.preprocess()
    dt = time.time() - t0
    out = cp.output_stream(); nt = cp.output_units()
    got = Result(int(cp.ok()), cp.n_vars(), out[0:2 * nt:2].copy(), out[2 * nt:].copy(), cp.undo(), {k: cp.stat(k) for k in STAT_NAMES})
    pb, rb = cp.stat("pair_batches"), cp.stat("resolve_batches")
    cp.close()
    want = run_oracle(stream, FULL)
    d = compare(got, want)
    bad += bool(d)
    print(f"rank {rank}/{world} {name}: {'OK' if not d else d[:3]}  {dt:.2f}s  eliminated {got.stats['bve_eliminatedVars']} pair_batches {pb} resolve_batches {rb}", flush=True)
t = torch.tensor([bad], device="cuda"); dist.all_reduce(t)
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_PARITY", "OK" if int(t) == 0 else "FAILED")
sys.exit(1 if int(t) else 0)

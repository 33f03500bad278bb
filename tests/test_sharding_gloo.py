"""N>1 host logic on CPU (gloo, world_size 2): the request partition and hit all-gather of the sharded scan
(gpu_db.cu cp3g_scan: contiguous slices `per = ceil(nreq / world)`, nccl_comm.cpp all-gather-v) produce the same
hit set on every rank as one unsharded scan; the variable partition of the BVE pair matrix (cp3g_bve_pairs) is
balanced and its gathered slices equal the unsharded matrix; and the max-over-ranks / sum-over-ranks reductions
bench.py uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

from harness import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["CP3_ROOT"]); sys.path.insert(0, os.path.join(os.environ["CP3_ROOT"], "tests"))
import numpy as np
import torch, torch.distributed as dist
import riss_solver_b200 as rs
from harness import load_gen

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
gen = load_gen()
lits, sizes = gen.random_ksat(3000, 12780, 3, seed=21, dup=0.04, sup=0.06, flip=0.06)
v = np.abs(lits).astype(np.int64) - 1
codes = np.sort((2 * v + (lits < 0)).astype(np.int32).reshape(-1, 3), axis=1).reshape(-1) if (sizes == 3).all() else None
if codes is None:
    out = (2 * v + (lits < 0)).astype(np.int32); ends = np.cumsum(sizes); st = ends - sizes
    for s, e in zip(st.tolist(), ends.tolist()): out[s:e] = np.sort(out[s:e])
    codes = out
svc = rs.ScanService(lib_path=os.environ["CP3_MOCK"])
svc.upload(int(np.abs(lits).max()), codes, sizes); svc.build()
ids = np.arange(sizes.size, dtype=np.uint32)
for kind in (rs.SUBSUME, rs.STRENGTHEN):
    # the slice rule of cp3g_scan
    per = (ids.size + world - 1) // world
    rb, re = min(ids.size, per * rank), min(ids.size, per * (rank + 1))
    mine, chk = svc.scan(kind, ids[rb:re])
    mine["req"] += rb                                   # slice-local request index -> global
    parts = [None] * world
    dist.all_gather_object(parts, (mine, chk))          # all-gather-v of the hit lists
    allh = np.concatenate([p[0] for p in parts]); total_chk = sum(p[1] for p in parts)
    allh = allh[np.lexsort((allh["clause"], allh["req"]))]
    full, full_chk = svc.scan(kind, ids)
    full = full[np.lexsort((full["clause"], full["req"]))]
    assert np.array_equal(allh, full), kind
    assert total_chk == full_chk
    # every rank must hold the identical merged list
    digest = torch.tensor([int(allh["clause"].astype(np.int64).sum()), allh.size], dtype=torch.int64)
    lo = digest.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    hi = digest.clone(); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
# K5 (BVE pair sizes): the variable range rule of cp3g_bve_pairs - contiguous ranges cut where the running
# number of (pos,neg) pairs crosses total*r/world - and the in-place all-gather-v of the size slices
off, refs = svc.download_occ()
n = int(np.abs(lits).max())
vs = np.arange(n, dtype=np.uint32)
pos = [refs[off[2 * x]:off[2 * x + 1]] for x in range(n)]
neg = [refs[off[2 * x + 1]:off[2 * x + 2]] for x in range(n)]
pair_off, full_sz = svc.bve_pairs(vs, pos, neg)
total = int(pair_off[-1])
cut = [0] + [int(np.searchsorted(pair_off, total * r // world, side="left")) for r in range(1, world)] + [n]
for r in range(1, world + 1):
    cut[r] = max(min(cut[r], n), cut[r - 1])
vb, ve = cut[rank], cut[rank + 1]
_, my_sz = svc.bve_pairs(vs[vb:ve], pos[vb:ve], neg[vb:ve])
parts = [None] * world
dist.all_gather_object(parts, my_sz)
assert np.array_equal(np.concatenate(parts), full_sz)
sizes_per_rank = [int(pair_off[cut[r + 1]] - pair_off[cut[r]]) for r in range(world)]
assert max(sizes_per_rank) <= total // world + int(np.diff(pair_off).max()) + 1     # balanced up to one variable
# the reductions bench.py reports: time = max over ranks, work = sum over ranks
t = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
w = torch.tensor([10.0], dtype=torch.float64); dist.all_reduce(w, op=dist.ReduceOp.SUM)
assert float(t) == float(world) and float(w) == 10.0 * world
dist.destroy_process_group()
print("RANK_OK", rank)
'''


def test_sharded_scan_world2_gloo(tmp_path):
    from fuzz_engine import build_mock
    mock = build_mock()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29591",
                   CP3_ROOT=ROOT, CP3_MOCK=mock)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o[-2000:]

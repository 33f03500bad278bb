"""Differential fuzz of the CUDA path (C ABI) against the CPU oracle: python scripts/fuzz_gpu.py [rounds] [seed]
(set CP3_PARSUB_MIN=0 CP3_STATIC_MIN=0 CP3_BULK_MIN=0 CP3_THREADS=3 to force the large-queue code paths, CP3_K2_SORT=1 the
radix-sort CSR build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from harness import run_oracle, compare, load_gen
from fuzz_engine import run_engine
from fuzz_oracle import OPTSETS, random_instance
gen = load_gen()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed); bad = 0
for r in range(rounds):
    lits, sizes = random_instance(rng); stream = gen.csr_to_stream(lits, sizes)
    opts = OPTSETS[int(rng.integers(0, len(OPTSETS)))]
    d = compare(run_engine(stream, opts), run_oracle(stream, opts))
    if d:
        bad += 1; print(r, opts, d[:3], flush=True)
        if bad > 5: break
print(f"FUZZ seed {seed}: {rounds} rounds, {bad} mismatches")
sys.exit(1 if bad else 0)

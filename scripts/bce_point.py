"""Stand-alone BCE (-bce) at BASELINE configs[1] size through the C ABI, against the sequential reference on the same formula.
usage: python scripts/bce_point.py [n_vars]"""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
OPTS = ["-enabled_cp3", "-bce", "-no-cp3_limited"]
lits, sizes = rs.gen.random_ksat(n, int(4.26 * n), 3, seed=12345); stream = rs.gen.csr_to_stream(lits, sizes)
for rep in range(2):
    cp = rs.Coprocessor(OPTS); cp.add_stream(stream)
    t = time.time(); cp.preprocess(); tp = time.time() - t
    st = {k: cp.stat(k) for k in ("bce_steps", "bce_testedLits", "bce_remBCE", "bce_batches", "bce_rows_ondemand", "gpu_bce_ns", "gpu_bce_bytes", "host_gpuwait_ns")}
    print(f"rep {rep}: preprocess {tp:.3f}s clauses_out={cp.n_clauses()}", st, flush=True)
    cp.close()
ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "refshim")
if os.path.exists(ref):
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.bin"); out = os.path.join(td, "o.txt"); stream.tofile(inp)
        subprocess.check_call([ref, "time", inp, out] + OPTS, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        print("reference sequential:", " ".join(l.strip() for l in open(out) if any(k in l for k in ("seconds", "bce_"))), flush=True)

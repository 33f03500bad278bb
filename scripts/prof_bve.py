import sys, time, os, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import riss_solver_b200 as rs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kind = sys.argv[2] if len(sys.argv) > 2 else "community"
opts = ["-enabled_cp3", "-up", "-subsimp", "-bve", "-no-cp3_limited"]
lits, sizes = rs.gen.community_ksat(n, 5 * n) if kind == "community" else rs.gen.random_ksat(n, int(4.26 * n), 3, seed=12345)
stream = rs.gen.csr_to_stream(lits, sizes)
for rep in range(2):
    cp = rs.Coprocessor(opts); cp.add_stream(stream)
    t = time.time(); cp.preprocess(); dt = time.time() - t
    names = ["sub_subsSteps", "sub_strengthSteps", "bve_steps", "bve_eliminatedVars", "bve_createdClauses", "scan_batches", "scan_ondemand", "pair_batches", "resolve_batches",
             "gpu_launches", "host_gpuwait_ns", "host_commit_ns", "ph_total_ns"]
    print(f"ours rep {rep}: {dt:.3f}s", {k: (round(cp.stat(k) / 1e6, 1) if k.endswith("_ns") else cp.stat(k)) for k in names})
    cp.close()
ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "refshim")
with tempfile.TemporaryDirectory() as td:
    inp = os.path.join(td, "in.bin"); out = os.path.join(td, "o.txt"); stream.tofile(inp)
    subprocess.check_call([ref, "time", inp, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print("reference sequential:", " ".join(l.strip() for l in open(out) if any(k in l for k in ("seconds", "bve_steps", "bve_eliminatedVars"))))

"""A/B of the staged device->host copies (csrc/staging.cu): CPaddLits (bulk ingest + arena download) and CPpreprocess at C2,
with CP3_NO_STAGED_D2H=1 (plain cudaMemcpy into pageable memory) and without, one process each."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import time
    sys.path.insert(0, root)
    import numpy as np, torch
    import riss_solver_b200 as rs
    from bench import Config
    cfg = Config(sys.argv[2], int(sys.argv[3]))
    lits, sizes, stream = cfg.generate()
    pinned = torch.from_numpy(np.ascontiguousarray(stream)).pin_memory().numpy()
    ta, tp, to = [], [], []
    for rep in range(6):
        cp = rs.Coprocessor(cfg.opts)
        t0 = time.perf_counter(); cp.add_stream(pinned); t1 = time.perf_counter(); cp.preprocess(); t2 = time.perf_counter(); out = cp.output_stream(); t3 = time.perf_counter()
        ta.append(t1 - t0); tp.append(t2 - t1); to.append(t3 - t2)
        chk = (cp.stat("sub_subsSteps"), cp.stat("sub_strengthSteps"), cp.n_clauses(), int(out.sum()))
        cp.close()
    med = lambda v: sorted(v[1:])[len(v[1:]) // 2] * 1e3
    print(f"{os.environ.get('TAG','')}: add {med(ta):.1f} ms  preprocess {med(tp):.1f} ms  output {med(to):.1f} ms  (medians of 5)  check {chk}", flush=True)
    sys.exit(0)
cfgname = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = sys.argv[2] if len(sys.argv) > 2 else "1000000"
for tag, env in (("plain ", {"CP3_NO_STAGED_D2H": "1"}), ("staged", {}), ("plain ", {"CP3_NO_STAGED_D2H": "1"}), ("staged", {})):
    e = dict(os.environ, TAG=tag, **env)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", cfgname, n], env=e, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-800:], flush=True)

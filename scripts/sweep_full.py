"""A/B of k_full build variants: K3/K4 launch times at C2, a C3-shaped point and two Zipf points per variant.
A variant is a library built from gpu_db.cu with other -D tile parameters (CP3_TILE, CP3_GROUPMAX, CP3_OFFCAP, CP3_WORKCAP, CP3_WQ, CP3_WH,
CP3_DCHUNK, CP3_MINBLOCKS) or a modified kernel, linked against the objects in csrc/_obj and stored as csrc/_obj/var/lib_<name>.so;
`name@B` runs it with CP3_FQ_BLOCKS=B blocks per SM.  Results of round 2: profiles/r02_kfull_sweep.txt."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1:]
for nm in names:
    lib, _, blocks = nm.partition("@")
    env = dict(os.environ, CP3B200_LIB=os.path.join(root, "riss-solver_b200", "csrc", "_obj", "var", f"lib_{lib}.so"))
    if blocks:
        env["CP3_FQ_BLOCKS"] = blocks
    for cmd in (["scripts/bench_kernels.py", "1000000", "4"], ["scripts/c3_kernels.py", "3000000"], ["scripts/bench_kernels.py", "1000000", "2", "1.0"], ["scripts/bench_kernels.py", "1000000", "2", "1.2"]):
        r = subprocess.run([sys.executable] + cmd, cwd=root, env=env, capture_output=True, text=True, timeout=900)
        print(f"== {nm} {' '.join(cmd)}\n" + "\n".join(r.stdout.strip().splitlines()[-2:]) + ("\n" + r.stderr[-400:] if r.returncode else ""), flush=True)

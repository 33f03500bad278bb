import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
import riss_solver_b200 as rs
n=int(sys.argv[1]); extra=sys.argv[2:]
lits,sizes=rs.gen.community_ksat(n,5*n)
stream=rs.gen.csr_to_stream(lits,sizes)
for rep in range(2):
    cp=rs.Coprocessor(["-enabled_cp3","-up","-subsimp","-bve","-no-cp3_limited"]+extra)
    cp.add_stream(stream); t=time.time(); cp.preprocess(); print("preprocess", time.time()-t)
    print({k:(round(cp.stat(k)/1e6,1) if k.endswith("_ns") else cp.stat(k)) for k in ["bve_eliminatedVars","pair_batches","resolve_batches","scan_batches","host_gpuwait_ns","host_commit_ns","ph_sub_ns","ph_str_ns","ph_init_ns","gpu_launches","arena_compactions"]})
    cp.close()

import sys, os, threading, time, ctypes, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
os.environ["ADTOMO_FORCE_CLUSTER"]="2"
import adtomo_jl_b200 as A, oracle, bench
L=A.load_library()
prog=torch.zeros(64,dtype=torch.int32).pin_memory()
torch.cuda.init(); torch.zeros(1,device='cuda')
L.adtomo_debug_set_progress.argtypes=[ctypes.c_void_p]
print("set", L.adtomo_debug_set_progress(prog.data_ptr()))
ctx=A.Context(0)
S=int(sys.argv[1])
w=bench.workload(1,0,s_per_gpu=S)
m,n,l=w['dims']; dims=(m,n,l)
ptr,idx,val=A.corner_sources(w['sta'],w['h'],w['vel0'])
u0=np.full((S,)+dims,1000.0)
for s in range(S): u0[s].ravel()[idx[ptr[s]:ptr[s+1]]]=val[ptr[s]:ptr[s+1]]
u=np.empty_like(u0); rounds=np.zeros(S,dtype=np.int32)
res={}
def run():
    res['rc']=ctx.forward3d_batch(u,u0,w['f'],w['h'],dims,1e-3,S,rounds=rounds)
th=threading.Thread(target=run,daemon=True); th.start()
for i in range(24):
    time.sleep(0.5)
    print(i, prog[:16].tolist(), flush=True)
    if not th.is_alive(): break
print("alive", th.is_alive(), rounds[:8], flush=True)
os._exit(0)

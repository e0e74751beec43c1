import sys, os, threading, time, ctypes, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
os.environ["ADTOMO_FORCE_CLUSTER"]="2"
import adtomo_jl_b200 as A, oracle
L=A.load_library()
prog=torch.zeros(64,dtype=torch.int32).pin_memory()
torch.cuda.init(); torch.zeros(1,device='cuda')
L.adtomo_debug_set_progress.argtypes=[ctypes.c_void_p]
print("set", L.adtomo_debug_set_progress(prog.data_ptr()))
ctx=A.Context(0)
dims=(12,10,8); S=int(sys.argv[1]) if len(sys.argv)>1 else 200
rng=np.random.default_rng(7)
f=0.5+rng.random(dims); u0=np.full((S,)+dims,1000.0)
for s in range(S): u0[s][tuple(rng.integers(0,d) for d in dims)]=0.0
u=np.empty_like(u0); rounds=np.zeros(S,dtype=np.int32)
res={}
def run():
    res['rc']=ctx.forward3d_batch(u,u0,f,0.3,dims,1e-6,S,rounds=rounds)
th=threading.Thread(target=run,daemon=True); th.start()
for i in range(12):
    time.sleep(0.5)
    print(i, prog[:16].tolist(), flush=True)
    if not th.is_alive(): break
print("alive", th.is_alive(), flush=True)
if not th.is_alive():
    ok=all(np.array_equal(u[s], oracle.eikonal3d_forward(u0[s],f,0.3,1e-6)[0]) for s in range(0,S,max(1,S//10)))
    print("equal", ok, rounds[:10])
os._exit(0)

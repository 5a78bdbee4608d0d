import ctypes, os, sys
sys.path.insert(0, '/root/repo')
import torch
from clipbert_b200 import _lib as L, ops
lib = L.lib(); dev="cuda"
def rnd(*s): return torch.randn(*s, device=dev).to(torch.bfloat16)
def timeit(**kw):
    for _ in range(3): ops.gemm(**kw)
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(50): ops.gemm(**kw)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return 1e3*e0.elapsed_time(e1)/50
for (M,N,K,bn) in [(1312,768,768,64),(1312,768,3072,64),(1312,768,3072,128),(1312,3072,768,256),(1312,2304,768,256),(12544,256,1024,64),(12544,256,1024,128),(50176,128,512,64),(50176,128,512,128),(8192,8192,2048,256)]:
    A,B = rnd(M,K), rnd(N,K); C = torch.zeros(M,N,device=dev,dtype=torch.bfloat16)
    base = dict(mode=0,m=M,n=N,k=K,a=A,a_rows=M,a_ld=K,b=B,b_rows=N,b_ld=K,out=C,out_ld=N,block_n=bn)
    res=[]
    for kch, cb in ((1,4),(1,2),(2,4),(2,2),(4,2),(0,0)):
        lib.cb_debug_gemm_kch(kch); lib.cb_debug_gemm_cbuf(cb)
        try: res.append("k%dc%d %.2f" % (kch, cb, timeit(**base)))
        except Exception as e: res.append("k%dc%d err" % (kch, cb))
    lib.cb_debug_gemm_kch(0); lib.cb_debug_gemm_cbuf(0)
    print("TN %dx%dx%d bn%d: %s us" % (M,N,K,bn," | ".join(res)), flush=True)
P,Mo,No = 1312,768,3072
dY,X = rnd(P,Mo), rnd(P,No); dW = torch.zeros(Mo,No,device=dev)
for bn in (128,256):
    res=[]
    for kch in (1,2):
        lib.cb_debug_gemm_kch(kch)
        if bn == 256 and kch == 2: continue
        res.append("kch%d %.2f" % (kch, timeit(mode=1,m=Mo,n=No,k=P,a=dY,a_rows=P,a_ld=Mo,b=X,b_rows=P,b_ld=No,split_k=1,out=dW,out_ld=No,out_fp32=1,block_n=bn)))
    lib.cb_debug_gemm_kch(0)
    print("WGRAD 768x3072x1312 bn%d: %s" % (bn, " | ".join(res)), flush=True)

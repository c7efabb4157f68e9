import torch, time
dev=torch.device('cuda')
def timeit(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    ts=[]
    flush=torch.zeros(256<<17,dtype=torch.float64,device=dev)
    for _ in range(n):
        flush.sum()
        e0.record(); f(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]*1e3
for mb in (8, 25.7, 57, 100, 400):
    n=int(mb*1e6/8)
    a=torch.randn(n,dtype=torch.float64,device=dev); b=torch.empty_like(a)
    t=timeit(lambda: b.copy_(a))
    print(f"copy {mb} MB read + {mb} MB write: {t:.1f} us  -> {2*mb/t*1e-3:.2f} TB/s")
    t=timeit(lambda: a.sum())
    print(f"   sum-read {mb} MB: {t:.1f} us -> {mb/t*1e-3:.2f} TB/s")
# gather: out[t] = src[idx[t]]
n=3_210_868; ns=32768*96
src=torch.randn(ns,dtype=torch.float64,device=dev)
idx_sorted=torch.sort(torch.randint(0,ns,(n,),device=dev))[0]
idx_rand=torch.randint(0,ns,(n,),device=dev)
out=torch.empty(n,dtype=torch.float64,device=dev)
for name,idx in (("sorted",idx_sorted),("random",idx_rand)):
    i32=idx.to(torch.int32)
    t=timeit(lambda: torch.index_select(src,0,i32,out=out))
    print(f"gather {name}: {t:.1f} us")

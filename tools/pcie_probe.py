import torch, time
n=398131200
a=torch.empty(n,dtype=torch.uint8).pin_memory(); b=torch.empty(n,dtype=torch.uint8).pin_memory()
da=torch.empty(n,dtype=torch.uint8,device='cuda'); db=torch.empty(n,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(h2d,d2h,reps=10):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): da.copy_(a,non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): b.copy_(db,non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/reps
run(1,1,2)
print("h2d only GB/s", n/run(1,0)/1e9); print("d2h only GB/s", n/run(0,1)/1e9); t=run(1,1); print("both: per-direction GB/s", n/t/1e9, "ms", t*1e3)

import torch, time
n=144*1024*1024
h_in=torch.empty(n,dtype=torch.uint8).pin_memory(); h_out=torch.empty(n,dtype=torch.uint8).pin_memory()
d_in=torch.empty(n,dtype=torch.uint8,device="cuda"); d_out=torch.empty(n,dtype=torch.uint8,device="cuda")
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,rep=5):
    best=1e9
    for _ in range(rep):
        torch.cuda.synchronize(); t0=time.perf_counter(); fn(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best*1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
def both(): h2d(); d2h()
print("h2d ms %.3f (%.1f GB/s)"%(t(h2d), n/t(h2d)/1e6)); print("d2h ms %.3f"%t(d2h)); print("both concurrently ms %.3f"%t(both))

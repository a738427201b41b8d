"""Raw pinned-memory copy bandwidth of the box (H2D, D2H, both at once): the ceiling of bench.py's end-to-end leg.  python tools/pcie_peak.py"""
import torch, time
dev=torch.device('cuda')
n=512*1024*1024
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
h2=torch.empty(n//2,dtype=torch.uint8).pin_memory(); d2=torch.empty(n//2,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(f,reps=5):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
a=t(lambda: d.copy_(h,non_blocking=True)); print('H2D alone GB/s', n/a/1e9)
b=t(lambda: h2.copy_(d2,non_blocking=True)); print('D2H alone GB/s', n/2/b/1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
c=t(both); print('duplex: H2D GB/s', n/c/1e9, 'D2H GB/s', n/2/c/1e9, 'time for 189MB up + 94MB down (ms)', c*189e6/n*1e3)

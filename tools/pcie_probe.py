import torch, time
n = 922*1024*1024//4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device='cuda')
m = 411*1024*1024//4
h2 = torch.empty(m, dtype=torch.float32).pin_memory()
d2 = torch.empty(m, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    h2.copy_(d2, non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); t3=time.perf_counter()
    print('H2D %.1f GB/s  D2H %.1f GB/s  both %.2f ms (H2D alone %.2f ms)'%(n*4/(t1-t0)/1e9, m*4/(t2-t1)/1e9, (t3-t2)*1e3, (t1-t0)*1e3))

import time, torch
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
for gb in (0.35, 1.5, 8, 24, 30):
    for rep in range(2):
        torch.cuda.empty_cache(); torch.cuda.synchronize()
        t0=time.perf_counter(); x=torch.empty(int(gb*2**30), dtype=torch.uint8, device="cuda"); torch.cuda.synchronize(); t1=time.perf_counter()
        del x; torch.cuda.synchronize(); t2=time.perf_counter(); torch.cuda.empty_cache(); torch.cuda.synchronize(); t3=time.perf_counter()
        print(gb, "malloc ms", round((t1-t0)*1e3,2), "free ms", round((t3-t2)*1e3,2))

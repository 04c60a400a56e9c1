import torch, time
n = 315*1024*1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record(); 
for _ in range(5): d.copy_(h, non_blocking=True)
e.record(); torch.cuda.synchronize()
print("H2D GB/s", 5*n/ (s.elapsed_time(e)*1e-3)/1e9)
s.record(); 
for _ in range(5): h.copy_(d, non_blocking=True)
e.record(); torch.cuda.synchronize()
print("D2H GB/s", 5*n/ (s.elapsed_time(e)*1e-3)/1e9)

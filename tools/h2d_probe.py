import torch, time
n = 177*1024*1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t = (time.perf_counter()-t0)/10
    print("H2D pinned GB/s", n/t/1e9)
# two concurrent streams
d2 = torch.empty(n, dtype=torch.uint8, device="cuda"); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); t = (time.perf_counter()-t0)/10
print("2-stream H2D GB/s", 2*n/t/1e9)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv", shell=True, capture_output=True, text=True).stdout)

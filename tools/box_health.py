"""Scratch: is this box's GPU healthy? copy bandwidth + bf16 matmul + clocks."""
import torch, time, subprocess
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): b.copy_(a)
e1.record(); torch.cuda.synchronize()
print("copy GB/s", 10 * 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3): x @ x
torch.cuda.synchronize(); e0.record()
for _ in range(10): x @ x
e1.record(); torch.cuda.synchronize()
print("bf16 TF/s", 10 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.mem,power.draw,utilization.gpu,memory.used", "--format=csv,noheader"], capture_output=True, text=True).stdout)
print(subprocess.run(["nvidia-smi", "--query-compute-apps=pid,used_memory", "--format=csv,noheader"], capture_output=True, text=True).stdout)

import torch, sys
sys.path.insert(0, "terra-gan_b200")
x = torch.empty(64*512*512*64, dtype=torch.bfloat16, device="cuda")
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"zero_ 2.1GB: {ms*1e3:.1f} us  {x.numel()*2/ms/1e6:.1f} GB/s")
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"copy_ 2.1GB: {ms*1e3:.1f} us  {2*x.numel()*2/ms/1e6:.1f} GB/s (r+w)")

import torch
x = torch.empty(1 << 30, dtype=torch.int32, device='cuda')   # 4 GiB
y = torch.empty(1 << 30, dtype=torch.int32, device='cuda')
def t(fn, nbytes, label, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f'{label}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s')
t(lambda: x.fill_(7), x.numel() * 4, 'fill (write-only 4 GiB)')
t(lambda: x.zero_(), x.numel() * 4, 'zero (memset 4 GiB)')
t(lambda: y.copy_(x), x.numel() * 8, 'copy (read+write 8 GiB)')
t(lambda: x.sum(), x.numel() * 4, 'sum (read-only 4 GiB)')

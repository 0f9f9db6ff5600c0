import torch
x=torch.empty(1<<30,dtype=torch.bfloat16,device='cuda')
y=torch.empty(1<<30,dtype=torch.bfloat16,device='cuda')
def t(f,n=5):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms=t(lambda: x.fill_(1.0)); print("fill 2GiB write-only: %.1f GB/s"%(2*(1<<30)/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy r+w: %.1f GB/s"%(4*(1<<30)/ms/1e6))
ms=t(lambda: x.sum()); print("sum read-only: %.1f GB/s"%(2*(1<<30)/ms/1e6))

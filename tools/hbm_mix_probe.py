"""What HBM bandwidth does this GPU reach for copy, write-only, read-only and a 5 read : 8 write mix (the pyramid build's)?
Used for DESIGN.md section 4.2: the pyramid writes 1.6x what it reads, and writes are the slower direction."""
import torch
dev=torch.device("cuda")
n=1<<30
a=torch.empty(n,dtype=torch.bfloat16,device=dev).normal_()
b=torch.empty_like(a)
def t(f,bytes_,reps=10):
    best=1e9
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return bytes_/best/1e6
print("copy   (1 read : 1 write): %.0f GB/s" % t(lambda: b.copy_(a), 2*a.numel()*2))
print("fill   (write only)      : %.0f GB/s" % t(lambda: b.fill_(1.0), a.numel()*2))
print("sum    (read only)       : %.0f GB/s" % t(lambda: a.sum(), a.numel()*2))
# 38 % read / 62 % write like the pyramid: read n/2.6 .. emulate: out[3n/5*..]: y = x.repeat? use expand-copy: write 1.6x what is read
x=a[: n*5//13]; y=b[: n*8//13].view(-1)
src=x.view(-1)
def mix():
    y[:src.numel()].copy_(src); y[src.numel():].fill_(0.5)   # two kernels; approximates the mix
print("mix    (5 read : 8 write, two kernels): %.0f GB/s" % t(mix, (src.numel()+y.numel())*2))

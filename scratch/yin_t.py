import sys, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
Y = torch.randn(64, 110335, device='cuda') * 0.1
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
print("yin 64x5s us", timeit(lambda: sb.core.yin(Y, 22050, 73.416, 587.33, 1024, 256)))

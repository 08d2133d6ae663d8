import sys, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for B in (1, 2, 4, 8):
    Y = torch.randn(B, 110335, device='cuda') * 0.1
    print(B, "x 5 s get_specs us", round(timeit(lambda: sb.transtacos_audio.get_specs(Y)), 2))

import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
from oracle import spectral_oracle as O
ys = [O.synth_speechlike(110335, 114514 + i) for i in range(4)]
f0s = sb.transtacos_audio.get_f0(ys)
for y, f in zip(ys, f0s):
    ref = O.tt_get_f0(y); r = np.abs(f / ref - 1)
    print("yin rel: median %.2e p99 %.2e frac<1e-4 %.4f frac<2e-3 %.4f" % (np.median(r), np.percentile(r, 99), (r < 1e-4).mean(), (r < 2e-3).mean()))
c0 = sb.transtacos_audio.get_c0(ys[0]); print("c0 max rel", np.max(np.abs(c0 / O.tt_get_c0(ys[0]) - 1)))
Y = torch.randn(64, 110335, device='cuda') * 0.1
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
print("frame_stats 64x5s us", timeit(lambda: sb.core.frame_stats(Y, 1024, 256)))
print("trim track 64x5s us", timeit(lambda: sb.core.frame_stats(Y, 512, 128, want_zcr=False)))
print("yin 64x5s us", timeit(lambda: sb.core.yin(Y, 22050, 73.416, 587.33, 1024, 256)))

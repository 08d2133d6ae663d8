"""Where the end-to-end step of bench.py spends host time: cProfile over get_specs(cpu_tensor[64, L], out=pinned)."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
ta = sb.transtacos_audio
B, L, T, F, M = 64, 110335, 431, 1025, 80
y = (torch.randn(B, L) * 0.1).pin_memory()
out = (torch.empty(B * T, F).pin_memory(), torch.empty(B * T, M).pin_memory())
for _ in range(5): ta.get_specs(y, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): ta.get_specs(y, out=out)
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): ta.get_specs(y, out=out)
pr.disable()
ps = pstats.Stats(pr); ps.sort_stats("tottime").print_stats(18)
for chunk in (8, 16, 24, 32):
    sc = ta.db_norm_scale(ta.hp)
    f = lambda: ta.features_host(ta.hp, y, 0.97, sc, sc, out=out, chunk=chunk)
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): f()
    torch.cuda.synchronize(); print("chunk", chunk, "ms", round((time.perf_counter() - t0) / 20 * 1e3, 3))

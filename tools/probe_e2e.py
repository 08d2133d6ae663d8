import sys, time, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
from transtacos_retunegan_b200 import transtacos_audio as ta, core
B, L, T, F, M = 64, 110335, 431, 1025, 80
dev = torch.device('cuda')
d = torch.empty(B * T * F, device=dev)
h = torch.empty(B * T * F, pin_memory=True)
for _ in range(3): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("D2H pinned GB/s", d.numel() * 4 / dt / 1e9, "ms", dt * 1e3)
hx = torch.empty(B * L, pin_memory=True); dx = torch.empty(B * L, device=dev)
t0 = time.perf_counter()
for _ in range(10): dx.copy_(hx, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("H2D pinned GB/s", hx.numel() * 4 / dt / 1e9)
y = (torch.randn(B, L) * 0.1).pin_memory()
mag = torch.empty(B * T, F).pin_memory(); mel = torch.empty(B * T, M).pin_memory()
cfg = ta.hp
sc = ta.db_norm_scale(cfg)
for chunk in (4, 8, 16, 32, 64):
    def run():
        ta.features_host(cfg, y, 0.97, sc, sc, out=(mag, mel), chunk=chunk)
    for _ in range(3): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("chunk", chunk, "ms", round(dt * 1e3, 3), "spec-s/s", round(B * L / 22050 / dt))

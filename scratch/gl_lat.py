import time, torch, sys
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
import bench
w = bench.make_griffinlim(sb, torch, 1, "tt")
for i in range(5): w.step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20): w.step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue ms/step", (t1 - t0) / 20 * 1e3, "total ms/step", (t2 - t0) / 20 * 1e3)
# graph
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    w.step(0)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        y = w.step(0)
torch.cuda.synchronize()
for i in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print("graph ms/step", e0.elapsed_time(e1) / 20)

import time, torch, sys
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
import bench
for specs in (False, True):
    w = bench.make_mstft(sb, torch, specs=specs, rot=1)
    for i in range(5): w.step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(50): w.step(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("specs", specs, "eager host ms/step", (t1 - t0) / 50 * 1e3, "total", (t2 - t0) / 50 * 1e3)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3): w.step(0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        grad = w.step(0)
    torch.cuda.synchronize()
    ref = grad.clone()
    for i in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("   graph ms/step", e0.elapsed_time(e1) / 50, "grad equal", torch.equal(ref, grad))

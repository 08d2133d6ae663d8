"""mstft step, eager against CUDA-graph replay: how much of the eager step is the host (Python + driver calls) and how much the
GPU.  The replayed graph is the whole step (forward, backward, the autograd wrapper's own ops); its gradient must equal the eager one."""
import sys, time
import torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
import bench
for specs in (False, True):
    w = bench.make_mstft(sb, torch, specs=specs, rot=1)
    for i in range(5): w.step(i)
    eager = w.step(0).clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(100): w.step(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3): w.step(0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        grad = w.step(0)
    for i in range(3): g.replay()
    torch.cuda.synchronize()
    same = torch.equal(eager, grad)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"specs={specs} eager: host {(t1 - t0) / 100 * 1e3:.4f} ms/step, total {(t2 - t0) / 100 * 1e3:.4f}; "
          f"graph replay {e0.elapsed_time(e1) / 100:.4f} ms/step; replayed gradient == eager: {same}")

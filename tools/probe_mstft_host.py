"""Where the host time of the eager mstft step goes (cProfile over 2000 steps; the step is host-bound: tools/probe_mstft_graph.py)."""
import cProfile, pstats, sys, io
import torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
import bench
specs = len(sys.argv) > 1 and sys.argv[1] == "specs"
w = bench.make_mstft(sb, torch, specs=specs, rot=1)
for i in range(20): w.step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(2000): w.step(i)
pr.disable()
torch.cuda.synchronize()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:6000])

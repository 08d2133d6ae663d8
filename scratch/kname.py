import torch, sys
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
from torch.profiler import profile, ProfilerActivity
y = torch.randn(4, 110335, device='cuda') * 0.1
sb.transtacos_audio.get_specs(y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sb.transtacos_audio.get_specs(y)
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(e.key[:100], e.device_time_total)

"""Under torchrun: multi_stft_loss(ddp_reduce=True) with the in-kernel exchange over NVLink peer memory against an NCCL all-reduce
of the rank-local losses (value must agree), and the step time of either path (SB200_DDP_PEER=0 selects NCCL)."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
import transtacos_retunegan_b200 as sb
B, T = 16, 22050
g = torch.Generator(device="cuda").manual_seed(100 + rank)
y = (0.1 * torch.randn(B, 1, T, device="cuda", generator=g)).clamp_(-0.999, 0.999)
yg = torch.tanh(1.1 * y).requires_grad_(True)
path = sb.loss.ddp_reduce_path()
for specs in (False, True):
    out = sb.multi_stft_loss(y, yg, ret_loss=True, ret_specs=specs, ddp_reduce=True)
    red = out[0] if specs else out
    loc = sb.multi_stft_loss(y, yg, ret_loss=True, ret_specs=specs)
    loc = (loc[0] if specs else loc).detach().clone()
    ref = loc.clone()
    dist.all_reduce(ref)
    ref /= world
    gathered = [torch.zeros_like(red) for _ in range(world)]
    dist.all_gather(gathered, red.detach())
    same = all(torch.equal(gathered[0], t) for t in gathered)
    (gr,) = torch.autograd.grad(red, yg)
    (gl,) = torch.autograd.grad((sb.multi_stft_loss(y, yg, ret_loss=True, ret_specs=specs)[0] if specs else sb.multi_stft_loss(y, yg, ret_loss=True)), yg)
    if rank == 0:
        print(f"[{path}] specs={specs} reduced {red.item():.9g} nccl-mean {ref.item():.9g} local {loc.item():.9g} "
              f"rel diff {abs(red.item() - ref.item()) / abs(ref.item()):.2e}; same on all ranks: {same}; grad == local grad: {torch.equal(gr, gl)}")
    def step():
        yg.grad = None
        o = sb.multi_stft_loss(y, yg, ret_loss=True, ret_specs=specs, ddp_reduce=True)
        (o[0] if specs else o).backward()
    for _ in range(20): step()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(300): step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda"); dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[{path}] specs={specs} step {dt.item() / 300 * 1e6:.1f} us (max over {world} ranks)")
dist.destroy_process_group()

"""C5 multi-GPU check: bmshj2018-hyperprior training step under DDP/NCCL, one process per GPU (torchrun).
Verifies that gradients are identical across ranks after the all-reduce and that the step runs on our kernels."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressai_environment_b200.zoo import bmshj2018_hyperprior
from compressai_environment_b200.training import RateDistortionLoss, configure_optimizers, train_step, wrap_ddp
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = bmshj2018_hyperprior(4).to(dev).train()
model = wrap_ddp(net, dev)
opt, aux_opt = configure_optimizers(net)
crit = RateDistortionLoss(lmbda=0.018)
g = torch.Generator(device=dev).manual_seed(100 + rank)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    x = torch.rand(16, 3, 256, 256, generator=g, device=dev)
    if it == 1: e0.record()
    out = train_step(model, crit, x, opt, aux_opt)
e1.record(); torch.cuda.synchronize()
w = net.g_a[0].weight.detach().flatten()[:1000].clone()
ws = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(ws, w)
same = all(torch.equal(ws[0], t) for t in ws)
if rank == 0:
    print(f"DDP world={world} loss={float(out['loss']):.4f} aux={float(out['aux_loss']):.2f} params_identical_across_ranks={same} "
          f"{16 * world * 3 / (e0.elapsed_time(e1) / 1e3):.1f} img/s")
assert same
dist.destroy_process_group()

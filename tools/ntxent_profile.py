import sys, torch, collections, re
sys.path.insert(0, '/root/repo')
from torch.profiler import profile, ProfilerActivity
from cstp_b200 import ops
rows, d, tau = 8192, 128, 0.1
z = torch.nn.functional.normalize(torch.randn(rows, d), dim=1).cuda()
lo, dz = torch.zeros(1, device='cuda'), torch.empty_like(z)
ws = torch.empty(ops.ntxent_workspace_floats(rows, d), device='cuda')
for _ in range(3): ops.ntxent(z, tau, True, lo, dz, ws)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.ntxent(z, tau, True, lo, dz, ws); torch.cuda.synchronize()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        print(f"{ev.device_time:9.1f} us  {ev.name[:90]}")

"""Time b200_zeldovich_ics_dev with and without the second-order term; --profile lists the kernels and
CUDA API calls of one 2LPT call (torch.profiler / CUPTI)."""
import sys, time; sys.path.insert(0, "lambda-cdm-raytracing_b200/python")
import torch, b200grav
eng = b200grav.Engine(0)
for g in (128, 256):
    n = g**3
    posm = torch.empty((n,4), device="cuda"); vel = torch.empty((n,3), device="cuda")
    for lpt in (0, 1):
        for rep in range(3):
            torch.cuda.synchronize(); t=time.perf_counter()
            st = eng.zeldovich_ics_dev(posm, vel, n_particles=n, grid=g, use_2lpt=lpt)
            torch.cuda.synchronize(); dt=time.perf_counter()-t
        print(f"grid {g} 2lpt={lpt}: {dt*1e3:.2f} ms rms={st[0]:.5f} max={st[1]:.5f}")
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        eng.zeldovich_ics_dev(posm, vel, n_particles=n, grid=g, use_2lpt=1)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))

import os, torch, torch.distributed as dist, time
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def bench(name, fn, nbytes):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0: print(f"PROBE {name}: {ms:.3f} ms  {nbytes/ms/1e6:.1f} GB/s (payload)", flush=True)
flat = torch.randn(19_000_000, device="cuda")
bench("all_reduce 76MB fp32", lambda: dist.all_reduce(flat), flat.numel()*4)
x = torch.randn(48, 162816, device="cuda").to(torch.bfloat16); out = torch.empty(48*world, 162816, device="cuda", dtype=torch.bfloat16)
bench("all_gather 15.6MB bf16", lambda: dist.all_gather_into_tensor(out, x), x.numel()*2)
big = torch.randn(3072*162816//4, device="cuda")
bench("all_reduce 500MB fp32", lambda: dist.all_reduce(big), big.numel()*4)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()

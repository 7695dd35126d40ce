"""2-GPU diagnostics: topology, peer access, raw NCCL send/recv bandwidth (torchrun)."""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print("peer access 0->1:", torch.cuda.can_device_access_peer(0, 1))
peer = rank ^ 1
for mib in (64, 1024, 4096):
    n = mib * (1 << 20) // 8
    a = torch.ones(n, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ops = [dist.P2POp(dist.isend, a, peer), dist.P2POp(dist.irecv, b, peer)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if rank == 0:
        print("sendrecv %d MiB each way: %.2f ms -> %.1f GB/s per direction" % (mib, dt * 1e3, mib * 1.048576e-3 / dt))
    del a, b
dist.destroy_process_group()

"""Step-by-step check of the peer-mapped exchange primitives (torchrun, N >= 2)."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nafwebsod_b200 as pkg
from nafwebsod_b200 import ops, _lib
from nafwebsod_b200.dp import _share_with_peers

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    torch.cuda.synchronize(dev); dist.barrier(); 
    if rank == 0: print(*a, flush=True)
buf = torch.full((1 << 20,), float(rank), device=dev)
flags = torch.zeros(64, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
peers, _ = _share_with_peers(buf, None)
pflags, _ = _share_with_peers(flags, None)
say("shared", [hex(p) for p in peers])
k = (rank + 1) % world
src = torch.full((1024,), 100.0 + rank, device=dev)
ops.p2p_copy(peers[k] + 4096, src.data_ptr(), 4096)
say("p2p_copy issued")
print(rank, "my buf[1024:1028] after peer wrote:", buf[1024:1028].tolist(), flush=True)
ops.p2p_signal([pflags[j] + 4 * rank for j in range(world)], 7)
say("signal issued")
ops.p2p_wait(flags[:world], 7, 5000, None)
say("wait done", flags[:world].tolist())
# bandwidth of copy-engine transfers
big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
pbig, _ = _share_with_peers(big, None)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.p2p_copy(pbig[k], big.data_ptr(), big.numel() * 4)
b.record(); torch.cuda.synchronize()
print(rank, "p2p copy 256 MiB: %.3f ms  %.0f GB/s" % (a.elapsed_time(b) / 10, big.numel() * 4 / (a.elapsed_time(b) / 10) / 1e6), flush=True)
dist.barrier()
dist.destroy_process_group()

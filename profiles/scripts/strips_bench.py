"""The row-strip leg of bench.py on its own (torchrun, one rank per GPU): python strips_bench.py [size] [levels]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import libdwt_b200 as d
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = d.lib(); L.init(local)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 0
def barrier():
    dist.barrier(); torch.cuda.synchronize()
peak, _ = bench.peaks()
r = bench.strips_leg(L, d, torch, dist, rank, world, peak, barrier, size=size, levels=levels)
if rank == 0:
    print(json.dumps(r))
dist.destroy_process_group()

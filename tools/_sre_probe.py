import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from lumol_b200 import _ffi, md, parallel, synthetic
from lumol_b200.device import DeviceSystem
rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
system = synthetic.lj_box((128, 128, 64), seed=20240 + 20)
synthetic.maxwell_boltzmann(system, 120.0, seed=7)
device = DeviceSystem(local_rank)
if world > 1:
    parallel.init_communicator(device, rank, world)
device.sync(system, velocities=True)
lib, ctx = device.lib, device.ctx
_ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, _ffi.INTEGRATOR_VELOCITY_VERLET, 1.0))
_ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 20))
_ffi.check(ctx, lib.lumol_cuda_reset_stats(ctx))
_ffi.check(ctx, lib.lumol_cuda_set_profiling(ctx, 1))
_ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 20))
_ffi.check(ctx, lib.lumol_cuda_set_profiling(ctx, 0))
st = device.stats()
print(f"rank {rank}: pair {st.pair_ms/20:.4f} integrate {st.integrate_ms/20:.4f} neighbor {st.neighbor_ms/20:.4f} comm(total) {st.comm_ms/20:.4f} halo {st.kspace_ms/20:.4f} ms/step", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

"""Developer probe (not product): on N ranks, time the halo exchange alone, pass 1 alone, and both posted together,
for the 6-D pair on a thin slab.  torchrun --nproc-per-node 2 tools/overlap_probe.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import levelsetpy_b200 as lsp
import bench
from levelsetpy_b200 import _lib as L
from levelsetpy_b200.slab import SlabSolver

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
gg, system, fill = bench.product_setup(lsp, "dubins6d", 41, 6 * world)
sol = SlabSolver(bench.scheme_for(lsp, gg, "as_shipped", system), device=local)
bench.fill_resident(sol.eng, gg, fill, slab=(sol.lo, sol.hi))
for _ in range(2):
    sol.step(0.0, 1e9, 0.8, L.COMP_MIN_OVER_TIME)
sol.begin_step(0.0, 1e9, 0.8)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ex_only():
    sol.comm.wait(sol.comm.post(*sol.halo_ops(0)))


def p1_only():
    sol.run_stage(1, which_pass=1)


def p2_only():
    sol.run_stage(1, which_pass=2)


def both():
    h = sol.comm.post(*sol.halo_ops(0))
    sol.run_stage(1, which_pass=1)
    sol.comm.wait(h)


def stage_ov():
    h = sol.comm.post(*sol.halo_ops(0))
    sol.run_stage(1, which_pass=1)
    sol.comm.wait(h)
    sol.finish_halos(0)
    sol.run_stage(1, which_pass=2)


def stage_serial():
    sol.comm.wait(sol.comm.post(*sol.halo_ops(0)))
    sol.finish_halos(0)
    sol.run_stage(1, which_pass=1)
    sol.run_stage(1, which_pass=2)


def fin_only():
    sol.finish_halos(0)


def step_full():
    sol.step(0.0, 1e9, 0.8, L.COMP_MIN_OVER_TIME)


r = dict(exchange=timed(ex_only), pass1=timed(p1_only), pass2=timed(p2_only), overlapped=timed(both),
         finish=timed(fin_only), stage_ov=timed(stage_ov), stage_serial=timed(stage_serial), step=timed(step_full))
if dist.get_rank() == 0:
    print("PROBE", {k: round(v, 3) for k, v in r.items()})
dist.barrier()
dist.destroy_process_group()

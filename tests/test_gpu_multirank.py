"""REAL multi-process slab solves: one process per GPU under ``torch.distributed.run`` (NCCL for the host-side group;
halos over peer memory and over NCCL send/recv), each slab compared with the single-domain answer bit for bit
(tests/slab_rank_worker.py).  Needs >= 2 GPUs on the box: on the one-GPU box the same halo code runs between contexts
of one process instead (tests/test_gpu_slab.py, LocalWorld)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_slab_ranks_match_single_domain(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "slab_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [json.loads(x) for x in r.stdout.splitlines() if x.startswith("{")]
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert len(lines) == 5 * 6, "every case x (transport, overlap, fused) must report"
    assert any(x["fused"] for x in lines) and any(x["pieces"] > 1 for x in lines)
    assert all(x["ok"] and x["bit_identical"] and x["dt_identical"] for x in lines)
    assert {x["protocol"] for x in lines} == {"two_pass", "ranged", "exchange_first"}

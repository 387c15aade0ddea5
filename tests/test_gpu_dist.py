"""Multi-GPU parity THROUGH THE DROP-IN CLASS: two processes (one per GPU) construct the model with the same full
arrays, exactly like the reference's single-process script (inference-adni-xval.py:127-133, MMM:1310), and must
reproduce the reference's golden traces — the class shards the trajectory axis, every EM iteration all-reduces the
sufficient statistics (peer-memory exchange), labels and (n_c, n) matrices are all-gathered.

Needs two visible GPUs; skipped otherwise (the same check runs inside `bench.py --gpus N` for N > 1)."""

import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.class_dist_checks import class_checks  # noqa: E402


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        os.environ["LOCAL_RANK"] = str(rank)
        import torch.distributed as dist

        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import mtm_b200  # noqa: F401
        from mtm_b200.framework import marginalizable_mixture_model as mix

        bad = class_checks(mix, world, rank)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, [f"exception: {e!r}\n{traceback.format_exc()}"]))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_class_shards_over_two_gpus():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(not bad for _, bad in res), res

"""world_size-2 gloo test of the hypothesis sharding used for multi-GPU runs (host tensors)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megapose6d_b200.parallel import HypothesisSharder


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = HypothesisSharder()
    full = torch.arange(n_rows * 3, dtype=torch.float32).view(n_rows, 3)
    s, e = sh.span(n_rows)
    out = sh.gather_rows(full[s:e] * 1.0, n_rows)
    q.put((rank, (s, e), torch.equal(out, full)))
    dist.destroy_process_group()


def _run(n_rows, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    return res


def test_gather_rows_even_and_ragged():
    for n_rows in (1152, 577, 1):
        res = _run(n_rows)
        spans = [r[1] for r in res]
        assert spans[0][0] == 0 and spans[-1][1] == n_rows and spans[0][1] == spans[1][0]
        assert all(r[2] for r in res), (n_rows, res)


def test_single_process_sharder_is_identity():
    sh = HypothesisSharder(enabled=False)
    assert sh.span(10) == (0, 10)
    t = torch.randn(4, 2)
    assert sh.gather_rows(t, 4) is t

"""Host-side logic of the data-parallel path with world_size 2 on CPU (gloo): shard bounds, global padded length and the
all-gather of finished mels (equal and unequal shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dexb200.parallel import gather_mels, global_max_length, padded_length, shard_bounds
        lo, hi = shard_bounds(n_total, rank, world)
        full = torch.arange(n_total * 80 * 8, dtype=torch.float32).reshape(n_total, 80, 8)
        y = full[lo:hi].clone()
        got = gather_mels(y, n_total=n_total)
        ok = torch.equal(got, full)
        if n_total % world == 0:
            ok = ok and torch.equal(gather_mels(y), full)
        lens = torch.tensor([37 + 10 * rank, 21])
        tmax = global_max_length(lens)
        ok = ok and tmax == 37 + 10 * (world - 1) and padded_length(tmax) == (tmax + 3) // 4 * 4
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_gather_world2_gloo(n_total):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.path.join(root, "dex-tts_b200") + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from dexb200.parallel import shard_bounds
    for n in (1, 7, 8, 64, 255, 256):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1

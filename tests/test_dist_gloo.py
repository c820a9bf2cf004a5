"""The N>1 plumbing on CPU: two gloo ranks shard distros by LPT, each fills its
padded result vector, one all-gather, every rank ends with the global vector."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from evergreen_b200 import _lib as L
from evergreen_b200 import dist as edist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = edist.lpt_partition(sizes, world)
    mine = shards.members[rank]
    local = np.zeros(shards.max_shard, dtype=L.ALLOC_RESULT_DTYPE)
    # what k_alloc would write for distro g: recognisable functions of the global id
    local["new_hosts"][:len(mine)] = mine * 3 + 1
    local["free_hosts"][:len(mine)] = mine % 7
    local["deficit_ns"][:len(mine)] = mine.astype(np.int64) * 10 ** 9
    send = torch.from_numpy(local.view(np.uint8).copy())
    got = edist.decode_results(edist.all_gather_results(send, shards, rank))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), got)
    # the double-buffered form used by bench.py: three ticks through two buffer sets
    pg = edist.PipelinedGather(shards, torch.device("cpu"))
    for k in range(3):
        pg.before_tick(k, None)
        buf = local.copy()
        buf["new_hosts"][:len(mine)] += k
        pg.send(k).copy_(torch.from_numpy(buf.view(np.uint8).copy()))
        pg.launch(k, None)
        assert np.array_equal(edist.decode_results(pg.result(k))["new_hosts"], np.arange(len(sizes)) * 3 + 1 + k)
    pg.drain(None)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_all_gather(tmp_path):
    sizes = np.array([5, 900, 30, 30, 1, 400, 2, 77, 12000], dtype=np.int64)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, sizes, str(tmp_path)), nprocs=2, join=True)
    ids = np.arange(len(sizes))
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"r{r}.npy"))
        assert np.array_equal(got["new_hosts"], ids * 3 + 1)
        assert np.array_equal(got["free_hosts"], ids % 7)
        assert np.array_equal(got["deficit_ns"], ids.astype(np.int64) * 10 ** 9)

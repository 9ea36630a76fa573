"""CPU, world_size 2 over gloo: the walker sharding / gather logic (the N>1 path).  The per-rank
evaluator is injected (the oracle on this rank's rows) because there is no GPU here; on the GPU
box the default evaluator is the CUDA path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, W, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as HH
    from artgpn_b200 import synth, sharding
    b = HH.cases()["batch"][0]
    d = HH.dataset(b["data"])
    theta = HH.arrays()[b["name"] + "_theta"][:W]

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()

    def evaluator(rows):
        vals = [HH.oracle.log_likelihood(d["time"], d["y"], d["yerr"], b["q"],
                                         *synth.objects_from_row(m, m, m, r, b["p"], b["q"], b["weights"])) for r in rows]
        return torch.tensor(vals, dtype=torch.float64)
    full = sharding.log_likelihood_batch_sharded(None, theta, evaluator=evaluator)
    np.save(out_path % rank, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W", [6, 5, 1])
def test_sharded_gather_world2(tmp_path, W):
    port = _free_port()
    out = str(tmp_path / "r%d.npy")
    mp.spawn(_worker, args=(2, port, W, out), nprocs=2, join=True)
    b = H.cases()["batch"][0]
    exp = H.arrays()[b["name"] + "_ll"][:W]
    r0, r1 = np.load(out % 0), np.load(out % 1)
    assert np.array_equal(r0, r1)                        # every rank holds the full vector
    assert r0.shape == (W,)
    assert H.rel_err(r0, exp) <= 1e-12


def test_shard_bounds_cover_everything():
    from artgpn_b200.sharding import shard_bounds, shard_sizes
    for W in (0, 1, 7, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(W, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == W
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert sum(shard_sizes(W, world)) == W
            assert max(shard_sizes(W, world)) - min(shard_sizes(W, world)) <= 1

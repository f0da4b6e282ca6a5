"""Multi-rank host logic on CPU: two `gloo` processes, each owning one aux block and one grid batch
(ProblemData.shard / dist.split_range), compute the partial sigma with the NumPy plan interpreter, combine it with
ONE all-reduce through `SigmaReducer` (the same object the GPU engine uses with NCCL), add the replicated local terms,
and must reproduce the unsharded oracle -- the N>1 contract of SURVEY 8(e).  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, method, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    from plan_interp import PlanInterpreter
    from xtddft_b200.dist import SigmaReducer, split_range
    from xtddft_b200.synth import make_problem
    from oracle.workloads import oracle_vind_for
    from xtddft_b200.workloads import plan_for
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = make_problem(13, 4, 2, 7, 11, 37, xctype="GGA", hyb=0.4, seed=321)     # naux=11, ng=37: ragged shards
        plan = plan_for(p, method)
        shard = p.shard(rank, world)
        p0, p1 = split_range(p.naux, rank, world)
        assert shard.cderi.shape[0] == p1 - p0
        it = PlanInterpreter(plan, shard)
        z = np.random.default_rng(5).standard_normal((3, plan.ext_dim))             # replicated trial vectors
        zs = it.pack(z)
        part = it.partial_blocks(zs)
        red = SigmaReducer()
        assert red.enabled and red.world == world and red.rank == rank
        flat = torch.from_numpy(np.concatenate([s.ravel() for s in part]))
        red.allreduce_(flat)                                                         # the one exchange step per vind call
        assert red.calls == 1 and red.bytes == flat.numel() * 8
        off, sig = 0, []
        for s in part:
            sig.append(flat[off:off + s.size].numpy().reshape(s.shape).copy())
            off += s.size
        it.add_local_blocks(zs, sig)
        got = it.unpack(sig)
        vind, hd = oracle_vind_for(p, method)
        ref = vind(z)
        err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
        # J-diagonal part of the XSF preconditioner is reduced the same way
        if plan.j_blocks:
            d = torch.from_numpy(it.jblock_diag(0).copy())
            red.allreduce_(d)
            full = PlanInterpreter(plan, p).jblock_diag(0)
            err = max(err, np.abs(d.numpy() - full).max())
        np.save(os.path.join(out_dir, f"err_{method}_{rank}.npy"), np.array([err]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("method", ["sf_down", "xtda", "xsf", "zvector"])
def test_two_rank_partial_sigma_allreduce(method, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, method, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err = float(np.load(tmp_path / f"err_{method}_{r}.npy")[0])
        assert err < 1e-11, (method, r, err)


def test_split_range_partitions():
    from xtddft_b200.dist import split_range
    for n in (0, 1, 7, 16, 4840):
        for world in (1, 2, 3, 8):
            parts = [split_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        split_range(4, 2, 2)

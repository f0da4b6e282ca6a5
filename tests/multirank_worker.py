"""Worker of tests/test_gpu_multirank.py: run under torchrun, one rank per GPU (NCCL).  Each rank owns one aux block
and one grid batch; sigma and the Davidson solve must equal the oracle / the single-rank result and be bit-identical
on every rank (replicated solver state)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist
    from oracle import sigma as osig
    from xtddft_b200 import davidson as pdav
    from xtddft_b200 import plan as planmod
    from xtddft_b200.dist import SigmaReducer, init_process_group_from_env
    from xtddft_b200.engine import SigmaEngine
    from xtddft_b200.synth import make_problem

    rank, local_rank, world = init_process_group_from_env("nccl")
    torch.cuda.set_device(local_rank)
    red = SigmaReducer()
    assert red.enabled and red.world == world
    p = make_problem(60, 12, 2, 46, 53, 1901, xctype="GGA", hyb=0.4, seed=11)          # naux, ng not divisible by world
    cases = [
        ("sf_down", planmod.build_sf_plan(p, isf=-1, method=0), osig.sf_gen_vind(p, -1, 0)),
        ("xtda", planmod.build_xtda_plan(p), osig.xtda_gen_vind(p)),
        ("xsf", planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf"),
         osig.xsf_gen_vind(p, sa=3, method=0, remove=True)),
    ]
    for name, plan, (vind, hd) in cases:
        eng = SigmaEngine.from_problem(plan, p, max_nvec=8, workspace_bytes=512 << 20, reducer=red, rank=rank, world=world)
        z = np.random.default_rng(3).standard_normal((5, hd.size))
        got = eng.sigma(torch.from_numpy(z).cuda())
        ref = vind(z)
        err = float(np.abs(got.cpu().numpy() - ref).max() / max(1.0, np.abs(ref).max()))
        assert err < 1e-9, (name, rank, err)
        assert np.abs(eng.hdiag() - hd).max() < 1e-10, name
        # bit-identical on every rank
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        for g in gathered:
            assert torch.equal(g, gathered[0]), name
        # Davidson: replicated state, same energies everywhere, equal to the dense spectrum of the oracle operator
        settings = {"sf_down": "sf_down", "xtda": "xtda", "xsf": "xsf"}[name]
        conv, e, x, cyc = pdav.davidson_for_engine(eng, 4, settings)
        a = np.asarray(vind(np.eye(hd.size)))
        w = np.linalg.eigvalsh(0.5 * (a + a.T))
        if name == "xtda":
            w = w[w > 1e-3]
        assert conv.all(), name
        assert np.abs(e - w[:4]).max() < 1e-6, (name, e, w[:4])          # north_star: energies within 1e-6 Eh
        et = torch.from_numpy(np.ascontiguousarray(e)).cuda()
        eg = [torch.empty_like(et) for _ in range(world)]
        dist.all_gather(eg, et)
        for g in eg:
            assert torch.equal(g, eg[0]), name
        if rank == 0:
            print(f"multirank {name}: world={world} sigma rel.err={err:.2e} davidson cycles={cyc[0]} ok", flush=True)
        eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

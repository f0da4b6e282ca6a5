"""TEST INFRASTRUCTURE (CPU oracle) -- oracle sigma builders matching the BASELINE workloads of
xtddft_b200/workloads.py::plan_for.  Imported by tests/ and by the cpu_baseline / --impl reference legs of bench.py only."""
from . import sigma as osig


def oracle_vind_for(p, method: str):
    if method == "xtda":
        return osig.xtda_gen_vind(p)
    if method == "sf_down":
        return osig.sf_gen_vind(p, -1, 0)
    if method == "sf_up":
        return osig.sf_gen_vind(p, 1, 0)
    if method == "xsf":
        return osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
    if method == "zvector":
        import numpy as np
        from . import zvector as ozv
        op = ozv.roks_matvec(p) if p.restricted else (lambda x, f=ozv.uks_fvind(p), g=ozv.uks_gaps(p): f(x) + g * np.asarray(x).ravel())
        return (lambda zs: np.stack([op(z) for z in np.atleast_2d(zs)])), None
    raise ValueError(method)

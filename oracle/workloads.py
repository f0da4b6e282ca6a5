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
    raise ValueError(method)

"""The stand-in mean-field objects that tests/golden/make_golden.py feeds to the REFERENCE's own classes, made available to the
adapter tests: `fake_scf()` installs the stub module tree (a fake `pyscf` etc.) for the duration of a `with` block and
yields (FakeROKS, FakeUKS); the stubs are removed from sys.modules afterwards so nothing else sees a fake PySCF."""
import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_make_golden():
    spec = importlib.util.spec_from_file_location("xtd_make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def fake_scf():
    before = set(sys.modules)
    mg = _load_make_golden()
    roks, uks = mg.install_stubs()
    try:
        yield roks, uks
    finally:
        for name in set(sys.modules) - before:
            if name.split(".")[0] in ("pyscf", "cupy", "gpu4pyscf", "opt_einsum", "pandas", "xtd_make_golden"):
                sys.modules.pop(name, None)


def packed_df(p, tensor="cderi", on_disk=False):
    """`mf.with_df` of a density-fitted PySCF object: lower-triangular packed rows, in memory or behind `loop()`."""
    full = getattr(p, tensor)
    il = np.tril_indices(p.nao)
    packed = np.ascontiguousarray(full[:, il[0], il[1]])
    if not on_disk:
        return types.SimpleNamespace(_cderi=packed, auxbasis="synthetic")

    def loop(blksize=7):
        for i in range(0, packed.shape[0], blksize):
            yield packed[i:i + blksize]
    return types.SimpleNamespace(_cderi="/nonexistent/cderi.h5", auxbasis="synthetic", loop=loop, get_naoaux=lambda: packed.shape[0])

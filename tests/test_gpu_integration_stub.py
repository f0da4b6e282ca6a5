"""The reference-side ctypes stub printed in INTEGRATION.md (Level 2) is executed verbatim against libxtdsigma.so and
must reproduce the oracle's SF-TDA `vind` -- so the documented binding is known to work with nothing but the C-ABI
(no plan compiler, no SigmaEngine).  Needs a B200."""
import os
import re
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    src = [b for b in blocks if "class B200Sigma" in b]
    assert len(src) == 1
    return src[0].replace('C.CDLL("libxtdsigma.so")', 'C.CDLL(LIBPATH)')


def test_integration_stub_matches_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import sigma as osig
    from xtddft_b200 import _lib
    from xtddft_b200.synth import make_problem
    p = make_problem(45, 9, 2, 34, 40, 900, xctype="LDA", hyb=0.5, seed=77, fxc_kinds=("alda0",))
    ns = {"LIBPATH": _lib.LIB_PATH}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    n = p.nao
    iu = np.tril_indices(n)
    mf = types.SimpleNamespace(
        mo_coeff=p.mo_coeff[0],
        mol=types.SimpleNamespace(nelec=(p.nocc_a, p.nocc_b)),
        with_df=types.SimpleNamespace(_cderi=np.ascontiguousarray(p.cderi[:, iu[0], iu[1]])))      # PySCF packed rows
    co = p.mo_coeff[0][:, :p.nocc_a]
    cv = p.mo_coeff[1][:, p.nocc_b:]
    fa_oo = p.fock_ks[0][:p.nocc_a, :p.nocc_a]
    fb_vv = p.fock_ks[1][p.nocc_b:, p.nocc_b:]
    op = ns["B200Sigma"](mf, co, cv, fa_oo, fb_vv, p.hyb, p.fxc_alda0, np.ascontiguousarray(p.ao[0]), p.weights)
    vind, hdiag = osig.sf_gen_vind(p, -1, 0)
    z = np.random.default_rng(3).standard_normal((4, hdiag.size))
    got, ref = op.vind(z), vind(z)
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    assert err < 1e-9, err          # north_star: sigma vectors within 1e-9 relative
    # a list of 1-D vectors is accepted like the reference's vind (SF_TDA.py:225-229)
    got1 = op.vind([z[0], z[1]])
    assert np.array_equal(got1, got[:2])

"""GPU: INTEGRATION.md section B, executed.  The ctypes stub printed there (`DeviceSweep`) is extracted from the
document, bound under the per-datum loop of the REFERENCE's own CRPMM / PCRPMM (oracle/_ref: the reference made
importable; the loop at pybgmm/igmm/crpmm.py:57-88 / pcrpmm.py:93-131 is replaced textually, the way a maintainer
would patch it), and the patched sampler must produce the unpatched reference's record_dict and assignments."""
import inspect
import os
import random
import re
import textwrap

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_namespace(lib_path):
    """The first python block of section B of INTEGRATION.md, executed against the built library."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## B."):]
    code = re.search(r"```python\n(.*?)```", sec, flags=re.S).group(1)
    assert "class DeviceSweep" in code
    ns = {}
    exec(compile(code.replace('"libbgmm_b200.so"', repr(lib_path)), "INTEGRATION.md#B", "exec"), ns)
    return ns


def _patched(cls, ns, first_line, last_marker, replacement):
    """`cls.collapsed_gibbs_sampler` with the source lines from `first_line` up to (not including) `last_marker`
    replaced by `replacement` -- the maintainer's patch of INTEGRATION.md section B."""
    src = textwrap.dedent(inspect.getsource(cls.collapsed_gibbs_sampler))
    lines = src.split("\n")
    a = next(i for i, ln in enumerate(lines) if ln.strip().startswith(first_line))
    b = next(i for i, ln in enumerate(lines) if i > a and ln.strip().startswith(last_marker))
    indent = re.match(r"\s*", lines[a]).group(0)
    lines[a:b] = [indent + ln for ln in replacement]
    loop = next(i for i, ln in enumerate(lines) if ln.strip().startswith("for i_iter in range(n_iter):"))
    lines.insert(loop, re.match(r"\s*", lines[loop]).group(0) + "dev = DeviceSweep(self.components)")
    glb = dict(inspect.getmodule(cls).__dict__)
    glb["DeviceSweep"] = ns["DeviceSweep"]
    exec(compile("\n".join(lines), "patched_" + cls.__name__, "exec"), glb)
    return type("Patched" + cls.__name__, (cls,), {"collapsed_gibbs_sampler": glb["collapsed_gibbs_sampler"]})


@pytest.mark.parametrize("which", ["G1_crpmm", "pcrpmm_r1.5"])
def test_stub_under_the_reference_loop(gpu_lib, which):
    try:
        from oracle.make_ref import import_ref
        NIW, CRPMM, PCRPMM, _, _ = import_ref()
    except Exception as e:  # the reference made importable travels with the snapshot; without it there is nothing to patch
        pytest.skip("oracle/_ref is not available here: %s" % e)
    ns = _stub_namespace(gpu_lib.LIB_PATH)
    if which == "G1_crpmm":     # pybgmm/tests/test_igmm.py:17-62
        N, D, K_true, seed, K, n_iter, kw = 100, 2, 4, 1, 3, 10, {}
        cls = _patched(CRPMM, ns, "for i in range(self.components.N):", "# Update record",
                       ["dev.sweep(self.alpha)", "dev.sync_back()"])
        ref_cls = CRPMM
    else:
        N, D, K_true, seed, K, n_iter, kw = 300, 3, 5, 7, 6, 6, dict(n_power=1.5, power_burnin=1)
        cls = _patched(PCRPMM, ns, "for i in data_loop_list:", "## end loop data",
                       ["dev.sweep(self.alpha, n_power if (flag_power and i_iter > power_burnin) else 1.0,",
                        "          None if isinstance(data_loop_list, range) else data_loop_list)",
                        "dev.sync_back()"])
        ref_cls = PCRPMM
    out = []
    for c in (ref_cls, cls):
        X, z_true = cases.gen(N, D, K_true, seed)
        m_0, k_0, v_0, S_0 = cases.prior_for(D, "full")
        model = c(X, NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments="rand", K=K, K_max=None)
        rec, _ = model.collapsed_gibbs_sampler(n_iter, z_true, num_saved=0, **kw)
        out.append((rec, model.components.assignments.copy(), model.components.K,
                    model.components.counts.copy(), model.log_marg(), random.random(), np.random.rand()))
    (rec_a, z_a, K_a, n_a, lm_a, r_a, nr_a), (rec_b, z_b, K_b, n_b, lm_b, r_b, nr_b) = out
    np.testing.assert_array_equal(z_b, z_a)
    assert K_b == K_a
    np.testing.assert_array_equal(n_b, n_a)
    assert (r_b, nr_b) == (r_a, nr_a), "the patched sampler must leave the global RNG streams where the reference does"
    for key in rec_a:
        if key == "sample_time":
            continue
        if key == "nk":
            assert rec_b[key] == rec_a[key]
        else:
            np.testing.assert_allclose(np.asarray(rec_b[key], dtype=float), np.asarray(rec_a[key], dtype=float),
                                       rtol=1e-9, err_msg=key)
    np.testing.assert_allclose(lm_b, lm_a, rtol=1e-9)
    if which == "G1_crpmm":
        assert z_b.tolist() == cases.G1_ASSIGNMENTS           # the reference's own golden vector

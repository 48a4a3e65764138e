"""Normal-inverse-Wishart hyper-parameters: the struct of `pybgmm/prior/niw.py:8-23` (fields m_0, k_0, v_0, S_0)."""


class NIW(object):
    """m_0: prior mean, shape (D,).  k_0: how strongly m_0 is believed (pseudo-count).  v_0: degrees of freedom --
    must be >= D (the reference asserts it, niw.py:21) and integer valued here (it indexes the lgamma table,
    gaussian_components.py:238).  S_0: D x D scale matrix, or a D-vector for the diagonal / NIX product model
    (gaussian_components_diag.py:92)."""

    def __init__(self, m_0, k_0, v_0, S_0):
        assert v_0 >= len(m_0), "v_0 must be larger or equal to dimension of data"
        self.m_0, self.k_0, self.v_0, self.S_0 = m_0, k_0, v_0, S_0

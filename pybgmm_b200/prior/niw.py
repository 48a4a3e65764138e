"""Normal-inverse-Wishart hyper-parameters (mirror of pybgmm/prior/niw.py:8-23: same fields, same assert)."""


class NIW(object):
    """m_0: prior mean (D,); k_0: belief in m_0; v_0: degrees of freedom (integer valued, >= D);
    S_0: D x D scale matrix (full covariance) or D-vector (diagonal / NIX product, gaussian_components_diag.py:92)."""

    def __init__(self, m_0, k_0, v_0, S_0):
        self.m_0 = m_0
        self.k_0 = k_0
        D = len(m_0)
        assert v_0 >= D, "v_0 must be larger or equal to dimension of data"
        self.v_0 = v_0
        self.S_0 = S_0

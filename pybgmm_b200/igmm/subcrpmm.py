"""Sub-clustering CRP mixture model (variable selection): the class surface of `pybgmm/igmm/subcrpmm.py:31-457`.

A binary mask over the D dimensions splits the data into the dimensions that are clustered (a CRP mixture over
X[:, mask == 1]) and the rest, explained by ONE common Gaussian component over X[:, mask == 0].  A sweep is the CRP sweep
of the masked data -- the same device path as CRPMM (`bgmm_sweep`, crpmm.py:57-88 == subcrpmm.py:400-431) -- followed,
after `burnin_mask` iterations, by a mask move: Gibbs over the dimensions (subcrpmm.py:306-337) or one Metropolis flip
(:192-291).  Each candidate mask is scored by log p(X, z | mask) + log p(mask), i.e. by `log_marg` of component objects
rebuilt on the candidate's columns with the current assignments (:293-304): on the device that is `bgmm_create` +
`bgmm_set_assignments` (the statistics kernel) + `bgmm_log_marg_k` per candidate.

The host policy follows the reference statement by statement, including its consumption of the global `numpy.random` and
`random` streams, so a seeded run is the reference's run.  Peculiarities kept on purpose: the Gibbs mask move never
rebuilds `self.components` (the clustering keeps running on the mask it started with; only the Metropolis move swaps the
component objects in, :253-256); `update_common_component` clears the last entry of the mask it is handed, in place, when
no dimension is left for the common component (:125-127); the partition term of a candidate reads the zero-count test from
`self.components` (:110).
"""
import copy
import logging
import math
import time

import numpy as np
from scipy.special import gammaln, logsumexp
from scipy.stats import bernoulli

from ..gaussian import GaussianComponents, GaussianComponentsDiag, GaussianComponentsFixedVar
from ..prior.betabern import BetaBern
from ..prior.niw import NIW
from ..utils import utils
from .igmm import IGMM

logger = logging.getLogger(__name__)

_CLASSES = {"full": GaussianComponents, "diag": GaussianComponentsDiag, "fixed": GaussianComponentsFixedVar}


class SubCRPMM(IGMM):

    def __init__(self, X, kernel_prior, alpha, save_path, assignments="rand", K=1, K_max=None, covariance_type="full",
                 common_component_covariance_type="full", bern_prior=BetaBern(1, 1), p_bern=0.1, device=0):
        # subcrpmm.py:35-36: the base class builds (and this class then replaces) components over all of X
        super(SubCRPMM, self).__init__(X, kernel_prior, alpha, save_path, assignments=assignments, K=K, K_max=K_max,
                                       covariance_type=covariance_type, device=device)
        self.components.chain.close()
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.covariance_type = covariance_type
        self.common_component_covariance_type = common_component_covariance_type
        self.K_max = K_max
        self.device = device
        self.h1 = 40   # prior information on the mean of the clustered dimensions (k_0 = 1 / h1)
        self.h0 = 40   # ... of the common component
        if bern_prior is not None:
            self.bern_prior = bern_prior
            self.p_bern = 1. * bern_prior.a / (bern_prior.a + bern_prior.b)
        else:
            self.bern_prior = None
            self.p_bern = p_bern
        self.make_robust_p_bern()
        self.total_run = 0
        self.update_run = 0
        self.acc_rate = 1
        # subcrpmm.py:62-63: every dimension starts included (the draw still comes from the global stream)
        self.mask = np.random.binomial(1, 1, self.D)
        self.common_component = self.update_common_component(1 - self.mask)
        if isinstance(assignments, str):
            # subcrpmm.py:72-88: a second, independent initialisation of the labels
            assignments = self._initial_assignments(assignments, K)
        self.components = self.update_clustering_components(self.mask, assignments)

    def make_robust_p_bern(self):
        """subcrpmm.py:93-103: an inclusion probability of exactly 0 or 1 would give -inf."""
        if self.p_bern == 1.0:
            self.p_bern = 1.0 - 0.0001
        if self.p_bern == 0.0:
            self.p_bern = 0.0001

    def log_marg_for_specific_component(self, components):
        """log p(X_mask, z) of a candidate's clustering components (subcrpmm.py:106-120)."""
        facts_ = gammaln(components.counts[:components.K])
        facts_[self.components.counts[:components.K] == 0] = 0
        log_prob_z = ((components.K - 1) * math.log(self.alpha) + gammaln(self.alpha)
                      - gammaln(np.sum(components.counts[:components.K]) + self.alpha) + np.sum(facts_))
        return log_prob_z + components.log_marg()

    @staticmethod
    def _data_prior(cols, h):
        """The data-driven NIW of subcrpmm.py:131-143 / :162-175: mean of the columns, k_0 = 1 / h, v_0 = D' + 2, S_0 = I."""
        d = cols.shape[1]
        return NIW(cols.mean(axis=0), 1.0 / h, d + 2, 1. * np.eye(d))

    def update_common_component(self, mask):
        """One component over the dimensions with mask == 0 (subcrpmm.py:122-154)."""
        common_D = np.where(mask == 0)[0].shape[0]
        if common_D == 0:
            mask[-1] = 0          # in place, like the reference
        common_X = self.X[:, np.where(mask == 0)[0]]
        prior = self._data_prior(common_X, self.h0)
        cls = _CLASSES.get(self.common_component_covariance_type)
        assert cls is not None, "Invalid covariance type."
        return cls(common_X, prior, np.zeros(common_X.shape[0]), 1, device=self.device)

    def update_clustering_components(self, mask, assignments):
        """The CRP mixture's components over the dimensions with mask == 1 (subcrpmm.py:156-185)."""
        cluster_X = self.X[:, np.where(mask == 1)[0]]
        if cluster_X.shape[1] == 0:
            raise ValueError("no dimension left for the clustering components")   # the reference fails in np.cov here
        prior = self._data_prior(cluster_X, self.h1)
        cls = _CLASSES.get(self.covariance_type)
        assert cls is not None, "Invalid covariance type."
        return cls(cluster_X, prior, assignments, self.K_max, device=self.device)

    @staticmethod
    def _release(*components):
        for c in components:
            c.chain.close()

    def log_marg_mask(self, mask_new):
        """log p(mask) + log p(X, z | mask) (subcrpmm.py:293-304)."""
        log_bern_new = np.sum(bernoulli.logpmf(mask_new, self.p_bern))
        common = self.update_common_component(mask_new)
        assert self.common_component.K == 1, "new common component can only have one cluster component"
        clustering = self.update_clustering_components(mask_new, self.components.assignments)
        value = self.log_marg_for_specific_component(clustering) + common.log_marg() + log_bern_new
        self._release(common, clustering)
        return value

    def gibbs_update_mask(self, i_iter):
        """subcrpmm.py:306-337: every dimension in turn, in or out by its conditional probability."""
        assert self.common_component.K == 1, "common component can only have one cluster component"
        for i_dim in range(self.D):
            mask_new = copy.deepcopy(self.mask)
            log_prob_mask = np.zeros(2, float)
            mask_new[i_dim] = 0
            log_prob_mask[0] = self.log_marg_mask(mask_new)
            mask_new[i_dim] = 1
            log_prob_mask[1] = self.log_marg_mask(mask_new)
            prob_mask = np.exp(log_prob_mask - logsumexp(log_prob_mask))
            self.mask[i_dim] = utils.draw(prob_mask)
        self.p_bern = np.random.beta(self.bern_prior.a + np.sum(self.mask),
                                     self.bern_prior.b + self.D - np.sum(self.mask), 1)[0]
        self.make_robust_p_bern()

    def metropolis_update_mask(self, i_iter):
        """subcrpmm.py:187-291: flip one dimension, accept by the ratio of log p(X, z, mask)."""
        assert self.common_component.K == 1, "common component can only have one cluster component"
        log_bern_old = np.sum(bernoulli.logpmf(self.mask, self.p_bern))
        log_marg_old = self.log_marg() + self.common_component.log_marg() + log_bern_old
        idx = np.random.choice(range(self.D), 1)[0]
        mask_new = copy.deepcopy(self.mask)
        mask_new[idx] = 1 - mask_new[idx]
        if np.where(mask_new == 1)[0].shape[0] == 0 or np.where(mask_new == 0)[0].shape[0] == 0:
            mask_new = copy.deepcopy(self.mask)
        log_bern_new = np.sum(bernoulli.logpmf(mask_new, self.p_bern))
        common = self.update_common_component(mask_new)
        assert self.common_component.K == 1, "new common component can only have one cluster component"
        clustering = self.update_clustering_components(mask_new, self.components.assignments)
        log_marg_new = self.log_marg_for_specific_component(clustering) + common.log_marg() + log_bern_new
        self.total_run += 1
        prob = np.exp(log_marg_new - log_marg_old)
        prob = 1 if prob > 1 else prob
        bern_prob = np.random.binomial(1, prob, 1)[0]
        if log_marg_new > log_marg_old or bern_prob > 0:
            self.update_run += 1
            self.acc_rate = self.update_run * 1. / self.total_run
            self._release(self.common_component, self.components)
            self.mask = copy.deepcopy(mask_new)
            self.common_component = common
            self.components = clustering
            if self.bern_prior is not None:
                self.p_bern = np.random.beta(self.bern_prior.a + np.sum(mask_new),
                                             self.bern_prior.b + self.D - np.sum(mask_new), 1)[0]
                self.make_robust_p_bern()
        else:
            self._release(common, clustering)

    def setup_subcrp_record(self):
        return {"included_variable": []}

    def update_subcrp_record(self, subcrp_record_dict):
        subcrp_record_dict["included_variable"].append(np.sum(self.mask))
        return subcrp_record_dict

    def collapsed_gibbs_sampler(self, n_iter, true_assignments, num_saved=3, weight_first=True, burnin_mask=500,
                                mask_update='gibbs', rng="reference"):
        """subcrpmm.py:352-457: `n_iter` iterations of (CRP sweep of the masked data, mask move once past `burnin_mask`);
        returns (record_dict, distribution_dict, subcrp_record_dict)."""
        record_dict = self.setup_record_dict()
        distribution_dict = self.setup_distribution_dict(num_saved)
        subcrp_record_dict = self.setup_subcrp_record()
        for i_iter in range(n_iter):
            if num_saved == self.components.K and i_iter > 1:
                distribution_dict = self.update_distribution_dict(distribution_dict, weight_first)
            tic = time.time()
            self._device_sweep(power=1.0, order=None, rng=rng)
            if i_iter > burnin_mask:
                if mask_update == 'gibbs':
                    self.gibbs_update_mask(i_iter)
                elif mask_update == 'metropolis':
                    self.metropolis_update_mask(i_iter)
                else:
                    assert False, "Invalid update method for mask vector."
            record_dict = self.update_record_dict(record_dict, i_iter, true_assignments, tic)
            subcrp_record_dict = self.update_subcrp_record(subcrp_record_dict)
        return record_dict, distribution_dict, subcrp_record_dict

"""Constrained-sampling CRP mixture model on the GPU engine.

API of `pybgmm/igmm/cscrpmm.py:21-485` (`CSCRPMM.constrained_gibbs_sample` and its convenience wrappers).  The sweep
itself -- including the constrained re-draw of cscrpmm.py:342-350 and the approximate step of :418-461 -- runs on the
device (`bgmm_sweep_constrained`); what the reference decides on the host per sweep is decided here the same way:
  * every `n_constrain`-th sweep is constrained: clusters with more than `thres * N` members are "useful", a datum of any
    other cluster is re-drawn until it lands in a useful one, a datum of a useful cluster stays (cscrpmm.py:155-167);
  * the count prior: log(n ** n_power) once `i_iter > power_burnin` (flag_power), or the per-sweep adaptive power
    1 + (r_up - 1) * (share of clusters with at most adapcrp_perct * N members) (flag_adapcrp_form2, :263-269), or the
    loss-driven power (flag_loss_adapcrp, :272-290);
  * the scan order is a fresh np.random.permutation(N) while the plain power is enabled (:292-297);
  * with flag_approx, sweeps after `approx_burnin` are followed by one constrained CRP sweep in data order (:418-461).
Every uniform comes from the interpreter's `random` stream, one per draw (utils.py:15), so the chain and the state the
generators are left in are the reference's.  Not on the device path (NotImplementedError): the deep-copy "loss" search
(flag_loss, :170-260, which draws from np.random inside the loop) and the per-datum adaptive power (flag_adapcrp, :307-314:
the power changes inside a sweep; `ADAPCRPMM` and flag_adapcrp_form2 are its per-sweep forms).
"""
import logging
import random
import time

import numpy as np

from .. import _lib
from ..utils import utils
from .igmm import IGMM

logger = logging.getLogger(__name__)


class CSCRPMM(IGMM):

    def _status(self, threshold):
        """1 = useful (more than `threshold` members), 2 = non-useful, per slot (cscrpmm.py:159-167)."""
        comps = self.components
        K = comps.K
        st = np.zeros(comps.K_max + 1, dtype=np.int32)
        st[:K] = np.where(comps.counts[:K] > threshold, 1, 2)
        return st

    def _constrained_sweep(self, power, order, status):
        """One constrained sweep: the uniforms are the next values of `random` -- one per datum plus one per re-draw, so
        a generous stretch is generated and the generator is then put where the draws actually made leave it."""
        comps = self.components
        state = random.getstate()
        budget = max(16 * comps.N, 65536) + comps.N
        u = _lib.mt19937_random(budget)
        try:
            st, used = comps.chain.sweep_constrained(self.alpha, power, order, u, status)
        except _lib.BgmmError as e:
            comps._dirty()
            if e.code == _lib.BGMM_EKMAX:
                raise IndexError(str(e))
            raise
        random.setstate(state)
        _lib.mt19937_random(used)
        comps._dirty()
        self.last_sweep_stats = st
        return st

    def constrained_gibbs_sample(self, n_iter, true_assignments,
                                 flag_constrain=False, n_constrain=1000000, thres=0.,
                                 flag_power=False, n_power=1, power_burnin=100000,
                                 flag_loss=False, n_loss_step=1000000, flag_marg=False, loss_burnin=10000000,
                                 flag_approx=False, approx_thres_perct=0., approx_burnin=1000000,
                                 flag_adapcrp=False, r_up=1., adapcrp_perct=0., adapcrp_burnin=1000000,
                                 flag_adapcrp_form2=False,
                                 flag_loss_adapcrp=False, r_up_losspcrp=1., lossadapcrp_step=0.,
                                 lossadapcrp_burnin=1000000,
                                 num_saved=3, weight_first=True):
        if flag_loss:
            raise NotImplementedError("flag_loss (deep-copy loss search, cscrpmm.py:170-260) is not on the device path")
        if flag_adapcrp:
            raise NotImplementedError("flag_adapcrp changes the power inside a sweep (cscrpmm.py:307-314); use "
                                      "flag_adapcrp_form2 or ADAPCRPMM (per-sweep power)")
        comps = self.components
        N = comps.N
        records = self.setup_record_dict()
        saved = self.setup_distribution_dict(num_saved)
        if flag_loss_adapcrp:
            smallest_loss = utils.cluster_loss_inertia(comps.X, comps.assignments)
            r_loss = 1.
        for i_iter in range(n_iter):
            if num_saved == comps.K and i_iter > 1:
                saved = self.update_distribution_dict(saved, weight_first)
            constrained = bool(flag_constrain) and i_iter % n_constrain == 0
            status = self._status(N * thres) if constrained else None
            if flag_adapcrp_form2 and i_iter > adapcrp_burnin:
                n_k = comps.counts[:comps.K]
                power_form2 = 1.0 + (r_up - 1.0) * (len(n_k[np.where(n_k <= N * adapcrp_perct)[0]]) * 1.0 / len(n_k))
            if flag_loss_adapcrp and i_iter > lossadapcrp_burnin:
                this_loss = utils.cluster_loss_inertia(comps.X, comps.assignments)
                if this_loss < smallest_loss:
                    r_loss -= lossadapcrp_step
                    smallest_loss = this_loss
                else:
                    r_loss += lossadapcrp_step
                r_loss = min(max(r_loss, 1.), r_up_losspcrp)
            order = np.random.permutation(N) if (flag_power and n_power > 1) else None
            if flag_power and i_iter > power_burnin:
                power = n_power
            elif flag_adapcrp_form2 and i_iter > adapcrp_burnin:
                power = power_form2
            elif flag_loss_adapcrp and i_iter > lossadapcrp_burnin:
                power = r_loss
            else:
                power = 1.0
            tic = time.time()
            if constrained:
                self._constrained_sweep(power, order, status)
            else:
                self._device_sweep(power=power, order=order)
            if flag_approx and i_iter > approx_burnin:
                self._constrained_sweep(1.0, None, self._status(N * approx_thres_perct))
            records = self.update_record_dict(records, i_iter, true_assignments, tic)
        return records, saved

    # the reference's convenience wrappers (cscrpmm.py:48-94)
    def approx_sampling(self, n_iter, _true_assignment, approx_thres_perct=0.04, approx_burnin=200, num_saved=3):
        return self.constrained_gibbs_sample(n_iter, _true_assignment, flag_approx=True,
                                             approx_thres_perct=approx_thres_perct, approx_burnin=approx_burnin,
                                             num_saved=num_saved)

    def ada_pcrp_sampling_form2(self, n_iter, _true_assignment, r_up=1.1, adapcrp_perct=0.04, adapcrp_burnin=500,
                                num_saved=3):
        return self.constrained_gibbs_sample(n_iter, _true_assignment, flag_adapcrp_form2=True, r_up=r_up,
                                             adapcrp_perct=adapcrp_perct, adapcrp_burnin=adapcrp_burnin,
                                             num_saved=num_saved)

    def constrained_sampling(self, n_iter, _true_assignment, n_constrain=10, thres=0.04, num_saved=3):
        return self.constrained_gibbs_sample(n_iter, _true_assignment, flag_constrain=True, n_constrain=n_constrain,
                                             thres=thres, num_saved=num_saved)

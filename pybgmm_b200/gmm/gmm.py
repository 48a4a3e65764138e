"""Per-sweep record bookkeeping shared by the samplers (the role of `pybgmm/gmm/gmm.py:28-118`)."""
import logging
import time

import numpy as np

from ..utils import utils
from ..utils.metrics import information_variation, mutual_information, normalized_mutual_information

logger = logging.getLogger(__name__)

#: keys of the record dict, in the reference's order (gmm.py:45-63)
RECORD_KEYS = ("sample_time", "log_marg", "components", "nmi", "mi", "nk", "loss", "bic", "vi", "alpha")


class GMM(object):
    """Record keys, their meaning and the every-20-sweeps log line are the reference's; what is new is
    `metrics_every`: the clustering metrics are O(N) host work that dominates once a sweep takes milliseconds."""

    #: evaluate nmi / mi / vi / loss every this many sweeps (1: every sweep, like the reference; 0: never).
    #: Sweeps without metrics record None for them.
    metrics_every = 1

    def label_switch(self, idx, nplist):
        return np.asarray(nplist)[idx]

    def setup_record_dict(self):
        return dict((key, []) for key in RECORD_KEYS)

    #: "device": the contingency table and the per-cluster squared distances come from the GPU (bgmm_contingency,
    #: bgmm_cluster_ssq; the labels stay there); "host": both are recounted from the labels with NumPy.
    metrics_backend = "device"

    def _clustering_metrics(self, true_assignments):
        comps = self.components
        table = None
        if self.metrics_backend == "device":
            table = comps.contingency(true_assignments)
            if table[:, -1].any():
                table = None          # unassigned data form a cluster of their own in the reference: count on the host
        if table is not None:
            table = table[:, :-1]
            z = None
            try:
                loss = utils.cluster_loss_from_ssq(comps.cluster_ssq())
            except NotImplementedError:   # fixed-variance statistics hold no sum of squares: count from the labels
                loss = utils.cluster_loss_inertia(comps.X, comps.assignments)
        else:
            z = comps.assignments
            loss = utils.cluster_loss_inertia(comps.X, z)
        return {"nmi": normalized_mutual_information(true_assignments, z, table=table),
                "mi": mutual_information(true_assignments, z, table=table),
                "vi": information_variation(true_assignments, z, base=2, table=table),
                "loss": loss,
                "bic": loss}   # the reference files the inertia loss under "bic" as well (gmm.py:99-101)

    def update_record_dict(self, record_dict, i_iter, true_assignments, start_time):
        """Append this sweep's row (gmm.py:65-118): the elapsed time is read first, so it covers the sweep only."""
        row = {"sample_time": time.time() - start_time,
               "log_marg": self.log_marg(),
               "components": self.components.K,
               "nk": str(self.components.counts[:self.components.K]),
               "alpha": self.alpha}
        wanted = (true_assignments is not None and bool(self.metrics_every)
                  and i_iter % self.metrics_every == 0)
        row.update(self._clustering_metrics(true_assignments) if wanted
                   else dict.fromkeys(("nmi", "mi", "vi", "loss", "bic")))
        for key in RECORD_KEYS:
            record_dict[key].append(row[key])
        if i_iter % 20 == 0:
            logger.info("iteration: %d, %s.", i_iter, ", ".join("%s: %s" % (k, row[k]) for k in sorted(row)))
        return record_dict

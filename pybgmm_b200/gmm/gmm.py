"""Base class with the per-sweep record bookkeeping (mirror of pybgmm/gmm/gmm.py:28-118)."""
import logging
import time

import numpy as np

from ..utils import utils
from ..utils.metrics import information_variation, mutual_information, normalized_mutual_information

logger = logging.getLogger(__name__)


class GMM(object):
    """Base class for the mixture samplers: record_dict keys and the every-20-sweeps log line are the reference's."""

    #: compute nmi / mi / vi / loss every `metrics_every` sweeps (1 = every sweep, as the reference; 0 = never,
    #: None is appended instead).  They are O(N) host work, which dominates once a sweep takes milliseconds.
    metrics_every = 1

    def __init__(self):
        pass

    def label_switch(self, idx, nplist):
        return np.array(nplist)[idx]

    def setup_record_dict(self):
        """gmm.py:45-63."""
        return {key: [] for key in ("sample_time", "log_marg", "components", "nmi", "mi", "nk", "loss", "bic", "vi",
                                    "alpha")}

    def update_record_dict(self, record_dict, i_iter, true_assignments, start_time):
        """gmm.py:65-118."""
        record_dict["sample_time"].append(time.time() - start_time)
        record_dict["log_marg"].append(self.log_marg())
        record_dict["components"].append(self.components.K)
        do_metrics = bool(self.metrics_every) and (i_iter % self.metrics_every == 0) and true_assignments is not None
        if do_metrics:
            z = self.components.assignments
            nmi = normalized_mutual_information(true_assignments, z)
            mi = mutual_information(true_assignments, z)
            loss = utils.cluster_loss_inertia(self.components.X, z)
            vi = information_variation(true_assignments, z, base=2)
        else:
            nmi = mi = loss = vi = None
        record_dict["nmi"].append(nmi)
        record_dict["mi"].append(mi)
        record_dict["nk"].append(str(self.components.counts[:self.components.K]))
        record_dict["loss"].append(loss)
        record_dict["bic"].append(loss)  # the reference stores the same quantity under "bic" (gmm.py:99-101)
        record_dict["vi"].append(vi)
        record_dict["alpha"].append(self.alpha)
        if i_iter % 20 == 0:
            info = "iteration: " + str(i_iter)
            for key in sorted(record_dict):
                info += ", " + key + ": " + str(record_dict[key][-1])
            info += "."
            logger.info(info)
        return record_dict

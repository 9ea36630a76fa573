"""Legacy network -- adapter for ``artgpn.art.network`` (art.py:12-354) on the same C ABI
(SURVEY.md section 8(f) #3).  One shared weight-kernel class whose amplitude is replaced per (series,
node) by ``weights_values[j + q*i]``; everything else goes through ``artgpn_b200.arttwo.network``.

Quirks of the reference that are reproduced on purpose:
  * the constructor stores ``yerr**2`` (art.py:58) and the likelihood squares it again (art.py:266), so
    sigma^4 ends up on the diagonal -- the inner network is simply fed ``yerr**2`` as its "sigma";
  * the weight class comes from the CONSTRUCTOR's ``weights[0]``, the other hyper-parameters from the
    ``weights[0]`` passed to the call, whose ``.pars[0]`` is overwritten in place (art.py:255-259);
  * ``prediction`` uses the constructor's jitter in the training block but the call's jitter in K**
    (art.py:320 vs :343) and returns ``(mean, std, cov)`` with the full M x M covariance (art.py:354).
"""
import numpy as np

from . import arttwo as _arttwo
from . import weight as _weight


class network(object):
    def __init__(self, nodes, weights, weights_values, means, jitters, time, *args):
        self.nodes = np.array(nodes)
        self.q = len(self.nodes)
        self.weights = np.array(weights)
        self.weights_values = np.array(weights_values)
        self.means = np.array(means)
        self.jitters = np.array(jitters)
        self.time = time
        self.args = args
        self.p = int(len(self.args) / 2)
        self.qp = self.q * self.p
        self.tt = np.tile(time, self.p)
        self.y = np.array([a for i, a in enumerate(args) if i % 2 == 0])
        self.yerr = np.array([np.asarray(a) ** 2 for i, a in enumerate(args) if i % 2 == 1])     # art.py:58
        assert self.means.size == self.p, 'The numbers of means should be equal to the number of components'
        assert len(args) / 2 == self.p, 'Given data and number of components dont match'
        inner_args = []
        for i in range(self.p):
            inner_args += [self.y[i], self.yerr[i]]
        self._net = _arttwo.network(self.q, time, *inner_args)

    # -- helpers ------------------------------------------------------------------------------------
    def _kernel_pars(self, kernel):
        return kernel.pars

    def _weights_list(self, weight, weights_values):
        """p*q weight objects of the constructor's class, amplitude from weights_values (art.py:253-261)."""
        cls = getattr(_weight, type(self.weights[0]).__name__)
        pars = weight.pars
        out = []
        for i in range(self.p):
            for j in range(self.q):
                pars[0] = weights_values[j + self.q * i]            # in-place, like the reference
                out.append(cls(*pars))
        return out

    def _kernel_matrix(self, kernel, time=None):
        return self._net._kernel_matrix(kernel, time)

    def _predict_kernel_matrix(self, kernel, tstar):
        return self._net._predict_kernel_matrix(kernel, tstar)

    def _mean(self, means, time=None):
        return self._net._mean(means, time)

    # -- reference API --------------------------------------------------------------------------------
    def _covariance_matrix(self, nodes, weight, weight_values, time, position_p, add_errors=False):
        """art.py:152-189"""
        return self._net._covariance_matrix(nodes, self._weights_list(weight, weight_values), time, position_p, add_errors)

    def compute_matrix(self, nodes, weight, weight_values, time, nugget=False, shift=False):
        """art.py:191-226 (the stored yerr**2 goes on the diagonal, no jitter)"""
        return self._net.compute_matrix(nodes, self._weights_list(weight, weight_values), time, nugget, shift)

    def log_likelihood(self, nodes, weights, weights_values, means, jitters):
        """art.py:228-276"""
        return self._net.log_likelihood(nodes, self._weights_list(weights[0], weights_values), means, jitters)

    def prediction(self, nodes=None, weights=None, weights_values=None, means=None, jitters=None, time=None, dataset=1):
        """art.py:280-354 -> (y_mean, y_std, y_cov)"""
        print('Working with dataset {0}'.format(dataset))
        nodes = nodes if nodes else self.nodes
        weights = weights if weights else self.weights
        weights_values = weights_values if weights_values else self.weights_values
        jitters = jitters if jitters else self.jitters
        return self._net.prediction_cov(nodes=nodes, weights=self._weights_list(weights[0], weights_values), means=means,
                                        jitters=jitters, time=time, dataset=dataset,
                                        train_jitter=self.jitters[dataset - 1])

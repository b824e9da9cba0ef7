"""Host-side helpers of the hot path's callers (mirror of the scalar / one-off parts of
gpdemo/utils.py).  These are O(1) / O(nD) host computations and stay on the host."""
import numpy as np
from scipy.special import gammaln


def gamma_log_pdf(x, a, b):
    """log Gamma(x; shape a, rate b)   (gpdemo/utils.py:19-36)"""
    return a * np.log(b) - gammaln(a) + (a - 1) * np.log(x) - b * x


def log_gamma_log_pdf(x, a, b):
    """log density of x = log(g), g ~ Gamma(shape a, rate b)   (gpdemo/utils.py:39-59).
    This is the theta prior every notebook closure adds to the log-ML estimate (nb cell 12)."""
    return a * np.log(b) - gammaln(a) + a * x - b * np.exp(x)


def adapt_factor_func(b, n_batch):
    """Adaptation schedule of the adaptive MH phase (gpdemo/utils.py:62-83): 5 -> 1.1 over the first
    fifth of the batches."""
    fifth = n_batch / 5.
    return 5. - min(b + 1, fifth) / fifth * 3.9


def normalise_inputs(X):
    """Zero-mean / unit-sd columns (gpdemo/utils.py:86-105); returns (X_norm, mean, sd)."""
    mean, sd = X.mean(0), X.std(0)
    return (X - mean[None]) / sd[None], mean, sd

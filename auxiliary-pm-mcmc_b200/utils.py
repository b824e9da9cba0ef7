"""Host-side helpers of the hot path's callers (mirror of the scalar / one-off parts of
gpdemo/utils.py).  These are O(1) / O(nD) host computations and stay on the host."""
import datetime
import json
import os

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


# ---- run artefacts: the file format of gpdemo/utils.py:108-208, so existing analysis notebooks read them ----------
def _perf_stats(n_reject, n_cubic_ops, comp_time):
    if hasattr(n_reject, '__len__'):
        return np.array([n for n in n_reject] + [n_cubic_ops, comp_time])
    return np.array([n_reject, n_cubic_ops, comp_time])


def _run_files(output_dir, tag):
    stamp = datetime.datetime.now().strftime('%Y_%m_%d_%H_%M_%S_')
    return (os.path.join(output_dir, stamp + tag + '_results.npz'), os.path.join(output_dir, stamp + tag + '_params.json'))


def save_run(output_dir, tag, thetas, n_reject, n_cubic_ops, comp_time, run_params):
    """<timestamp><tag>_results.npz with `thetas` and `n_reject_n_cubic_ops_comp_time`, and <timestamp><tag>_params.json
    (sorted keys, indent 4) -- gpdemo/utils.py:108-149.  Returns the two paths."""
    results_file, params_file = _run_files(output_dir, tag)
    np.savez(results_file, thetas=thetas, n_reject_n_cubic_ops_comp_time=_perf_stats(n_reject, n_cubic_ops, comp_time))
    with open(params_file, 'w') as f:
        json.dump(run_params, f, indent=4, sort_keys=True)
    return results_file, params_file


def save_adaptive_run(output_dir, tag, adapt_thetas, adapt_prop_scales, adapt_accept_rates, thetas, n_reject,
                      n_cubic_ops, comp_time, run_params):
    """As save_run plus the three arrays of the adaptive phase (gpdemo/utils.py:152-208)."""
    results_file, params_file = _run_files(output_dir, tag)
    np.savez(results_file, adapt_thetas=adapt_thetas, adapt_prop_scales=adapt_prop_scales,
             adapt_accept_rates=adapt_accept_rates, thetas=thetas,
             n_reject_n_cubic_ops_comp_time=_perf_stats(n_reject, n_cubic_ops, comp_time))
    with open(params_file, 'w') as f:
        json.dump(run_params, f, indent=4, sort_keys=True)
    return results_file, params_file


# ---- chain diagnostics (the notebooks call R's coda through rpy2 for these; numpy here) ----------------------------
def effective_sample_size(x):
    """Effective sample size of one scalar chain, coda::effectiveSize style: n var(x) / spectral density at
    frequency 0, the spectrum estimated from an AR(p) fit with p chosen by AIC (Yule-Walker, order <= 10 log10 n)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    xc = x - x.mean()
    var = xc.dot(xc) / n
    if n < 4 or var == 0.:
        return float(n) if var > 0. else 0.
    max_order = min(n - 1, int(np.floor(10. * np.log10(n))))
    acov = np.array([xc[:n - k].dot(xc[k:]) / n for k in range(max_order + 1)])
    best_aic, best_spec0 = None, var
    phi = np.zeros(0)
    v = acov[0]
    for p in range(0, max_order + 1):
        if p > 0:                                   # Levinson-Durbin step
            k = (acov[p] - phi.dot(acov[p - 1:0:-1] if p > 1 else np.zeros(0))) / v
            phi = np.r_[phi - k * phi[::-1], k]
            v = v * (1. - k * k)
            if not v > 0.:
                break
        aic = n * np.log(v) + 2. * (p + 1)
        if best_aic is None or aic < best_aic:
            best_aic = aic
            best_spec0 = (v * n / max(n - (p + 1), 1)) / (1. - phi.sum())**2
    return float(n * var / best_spec0) if best_spec0 > 0. else float(n)


def gelman_rubin(chains):
    """Potential scale reduction factor R-hat of m scalar chains of equal length (coda::gelman.diag point
    estimate without the degrees-of-freedom correction): sqrt(((n-1)/n W + B/n) / W)."""
    c = np.asarray(chains, dtype=np.float64)
    m, n = c.shape
    W = c.var(axis=1, ddof=1).mean()
    B = n * c.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((n - 1.) / n * W + B / n) / W))

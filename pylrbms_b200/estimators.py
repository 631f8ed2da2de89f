"""The a posteriori error estimator of the LRBMS (reference ``estimators.py:26-136``).

``EllipticEstimator.estimate(U, mu, d, decompose)`` keeps the reference signature.  Two execution paths, both GPU:

* ``d`` is a :class:`pylrbms_b200.reduced.ReducedModel` -- the online hot path.  All subdomains, all quadratic
  forms and the eta combine run in the mu-batched kernels behind ``lrbms_online_estimate`` (K5); ``U`` may hold
  any number of reduced solutions, one parameter each.
* ``d`` is a fine-scale discretization -- the reference's own sequence of ``pairwise_apply2`` calls
  (``estimators.py:70-91``), executed with SpMM / dot kernels, one parameter per call.

Reference quirks kept on purpose: ``nc`` / ``r`` / ``df`` are squared quantities
that are *not* square-rooted before the l2 norm over subdomains; the indicators square them again; the Poincare
constant is ``1/pi^2``; ``alpha`` returns inside its loop, i.e. only ``theta_0`` is looked at
(``alpha_returns_first=True``, the default).  One deviation, needed for batching: the norm over subdomains is taken
*per parameter column* (the reference's ``mpi_norm`` of an ``(S, len(U))`` array is one Frobenius scalar, which
is only meaningful for ``len(U) == 1``; for ``len(U) == 1`` both agree).
"""
from __future__ import annotations

import copy

import numpy as np

from .parameters import evaluate as _ev


class EllipticEstimator:
    def __init__(self, subdomains_on_rank, min_diffusion_evs, subdomain_diameters, local_eta_rf_squared,
                 lambda_coeffs, mu_bar, mu_hat, flux_reconstruction, oswald_interpolation_error,
                 alpha_returns_first=True, mpi_comm=None):
        self.subdomains = list(subdomains_on_rank)
        self.min_diffusion_evs = np.asarray(min_diffusion_evs, dtype=np.float64)
        self.subdomain_diameters = np.asarray(subdomain_diameters, dtype=np.float64)
        self.local_eta_rf_squared = np.asarray(local_eta_rf_squared, dtype=np.float64)
        self.lambda_coeffs = list(lambda_coeffs)
        self.mu_bar, self.mu_hat = mu_bar, mu_hat
        self.flux_reconstruction = flux_reconstruction
        self.oswald_interpolation_error = oswald_interpolation_error
        self.num_subdomains = len(self.subdomains)
        self.alpha_returns_first = bool(alpha_returns_first)
        self.mpi_comm = mpi_comm

    def with_(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    # -- reference estimators.py:114-130
    def alpha(self, thetas, mu, mu_bar):
        result = np.inf
        for theta in thetas:
            theta_mu, theta_mu_bar = _ev(theta, mu), _ev(theta, mu_bar)
            assert theta_mu / theta_mu_bar > 0
            result = min(result, theta_mu / theta_mu_bar)
            if self.alpha_returns_first:
                return result                       # reference estimators.py:121 returns inside the loop
        return result

    def gamma(self, thetas, mu, mu_bar):
        result = -np.inf
        for theta in thetas:
            theta_mu, theta_mu_bar = _ev(theta, mu), _ev(theta, mu_bar)
            assert theta_mu / theta_mu_bar > 0
            result = max(result, theta_mu / theta_mu_bar)
        return result

    def r_scale(self):
        """``(1/pi^2) / min_diffusion_ev * h^2`` per subdomain (reference ``estimators.py:88-91``)."""
        return (1.0 / np.pi ** 2) / self.min_diffusion_evs * self.subdomain_diameters ** 2

    def estimate(self, U, mu, d, decompose=False):
        if hasattr(d, 'estimate_batch'):
            # online hot path: one kernel sweep for all subdomains (and all parameters, if several are given)
            return d._estimate_with(self, U, mu, decompose)
        return self._estimate_elliptic_generic(U, mu, d, decompose)

    def _estimate_elliptic_generic(self, U, mu, d, decompose=False):
        """reference ``estimators.py:45-112`` on fine-scale arrays (generic operator chain)."""
        alpha_mu_mu_bar = self.alpha(self.lambda_coeffs, mu, self.mu_bar)
        gamma_mu_mu_bar = self.gamma(self.lambda_coeffs, mu, self.mu_bar)
        alpha_mu_mu_hat = self.alpha(self.lambda_coeffs, mu, self.mu_hat)
        n = len(U)
        local_eta_nc = np.zeros((self.num_subdomains, n))
        local_eta_r = np.zeros((self.num_subdomains, n))
        local_eta_df = np.zeros((self.num_subdomains, n))
        U_r = self.flux_reconstruction.apply(U, mu=mu)
        U_o = self.oswald_interpolation_error.apply(U)
        scale = self.r_scale()
        for ii, subdomain in enumerate(self.subdomains):
            local_eta_nc[ii] = d.operators['nc_{}'.format(subdomain)].pairwise_apply2(U_o, U_o, mu=mu)
            local_eta_r[ii] += self.local_eta_rf_squared[ii]
            local_eta_r[ii] -= 2 * d.operators['r_fd_{}'.format(subdomain)].apply(U_r, mu=mu).data[:, 0]
            local_eta_r[ii] += d.operators['r_dd_{}'.format(subdomain)].pairwise_apply2(U_r, U_r, mu=mu)
            local_eta_df[ii] += d.operators['df_aa_{}'.format(subdomain)].pairwise_apply2(U, U, mu=mu)
            local_eta_df[ii] += d.operators['df_bb_{}'.format(subdomain)].pairwise_apply2(U_r, U_r, mu=mu)
            local_eta_df[ii] += 2 * d.operators['df_ab_{}'.format(subdomain)].pairwise_apply2(U, U_r, mu=mu)
            local_eta_r[ii] *= scale[ii]
        eta = (np.sqrt(gamma_mu_mu_bar) * np.linalg.norm(local_eta_nc, axis=0)
               + np.linalg.norm(local_eta_r + local_eta_df, axis=0) / np.sqrt(alpha_mu_mu_hat)) / np.sqrt(alpha_mu_mu_bar)
        if n == 1:
            eta = float(eta[0])
        if decompose:
            local_indicators = (2.0 / alpha_mu_mu_bar) * (gamma_mu_mu_bar * local_eta_nc ** 2
                                                          + (1.0 / alpha_mu_mu_hat) * (local_eta_r + local_eta_df) ** 2)
            return eta, (local_eta_nc, local_eta_r, local_eta_df), local_indicators
        return eta

"""Online adaptive enrichment: Doerfler marking and the enrichment loop around the LRBMS hot path.

Drop-in for the reference's ``online_enrichment.py`` (``doerfler_marking`` ``:9-22``, ``AdaptiveEnrichment`` ``:25-93``):
same constructor arguments, same ``solve(mu, enrichment_steps, callback)`` return value ``(U, rd, reductor)``, same
marking rules (Doerfler on the *squared* indicators, plus every subdomain whose age exceeds ``marking_max_age``),
same age bookkeeping -- organised around a parameter batch: one ``rd.sweep`` (``lrbms_online_sweep``) per pass, then
``reductor.enrich_local`` -> ``d.solve_for_local_correction`` (``lrbms_pcg_solve``) -> ``extend_basis_local``
(``lrbms_va_*``), and ``reductor.reduce()`` (the batched projection plans).
"""
from __future__ import annotations

import logging

import numpy as np


def doerfler_marking(indicators, theta):
    """reference ``online_enrichment.py:9-22``: smallest set of largest *squared* indicators whose sum exceeds
    ``theta`` times the total (ties keep the reference's stable descending sort)."""
    assert 0.0 < theta <= 1.0
    indicators = np.asarray(indicators, dtype=np.float64).ravel() ** 2
    order = sorted(range(len(indicators)), key=lambda i: indicators[i], reverse=True)
    vals = indicators[order]
    total = np.sum(vals)
    sums = np.array([np.sum(vals[:ii + 1]) for ii in range(len(vals))])
    where = sums > theta * total
    if np.any(where):
        return [int(i) for i in order[:int(np.argmax(where)) + 1]]
    return [int(i) for i in order]


class AdaptiveEnrichment:
    """Estimator-driven enrichment of the local bases (the role of reference ``online_enrichment.py:25-93``).

    Built around a parameter BATCH: every pass is one ``rd.sweep`` over all parameters (solve + estimate + indicators in
    one launch set), the parameter with the largest estimate drives the enrichment -- its indicators are marked (Doerfler
    on the squared indicators plus every subdomain older than ``marking_max_age``, the reference's rules), its solution
    feeds ``reductor.enrich_local`` on the marked subdomains (``lrbms_pcg_solve`` + Gram-Schmidt), and the model is
    re-reduced (incrementally, if the reductor is set up for it).  ``solve(mu, ...)`` -- the reference's entry point, same
    arguments, same ``(U, rd, reductor)`` result, same callback dictionary -- is the batch of one."""

    def __init__(self, grid_and_problem_data, discretization, block_space, reductor, rd,
                 target_error, marking_doerfler_theta, marking_max_age):
        self.grid_and_problem_data = grid_and_problem_data
        self.discretization = discretization
        self.block_space = block_space
        self.reductor = reductor
        self.rd = rd
        self.target_error = target_error
        self.marking_doerfler_theta = marking_doerfler_theta
        self.marking_max_age = marking_max_age
        self.logger = logging.getLogger('pylrbms_b200.AdaptiveEnrichment')

    @property
    def _num_blocks(self):
        bs = self.block_space
        return bs.num_blocks if hasattr(bs, 'num_blocks') else len(bs.subspaces)

    def estimate(self, U, mu, decompose=False):
        return self.rd.estimate(U, mu=mu, decompose=decompose)

    def mark(self, indicators, ages):
        """Subdomains to enrich, ascending (the reference walks an unordered set): Doerfler marking of ``indicators`` united
        with every subdomain whose age exceeds ``marking_max_age``."""
        by_estimate = doerfler_marking(indicators, self.marking_doerfler_theta)
        by_age = np.flatnonzero(np.asarray(ages) > self.marking_max_age)
        self.logger.info('marked %d/%d subdomains by the estimator, %d more by age', len(by_estimate), self._num_blocks,
                         len(set(by_age.tolist()) - set(by_estimate)))
        return sorted(set(by_estimate) | {int(i) for i in by_age})

    def _enrich(self, U, mu, indicators, ages):
        """One enrichment of the model from the solution ``U`` (length 1) at ``mu``; returns the number of corrector solves."""
        marked = self.mark(indicators, ages)
        for subdomain in marked:
            self.reductor.enrich_local(subdomain, U, mu)
        before = self.rd.solution_space.dim
        self.rd = self.reductor.reduce()
        ages += 1
        ages[marked] = 1
        self.logger.info('%d corrector solves, reduced system %d -> %d', len(marked), before, self.rd.solution_space.dim)
        return len(marked)

    def _passes(self, mus, enrichment_steps, report, stop_when_exhausted):
        """The loop shared by ``solve`` and ``solve_batch``: sweep, report, stop or enrich from the worst parameter."""
        from .reductor import ExtensionError
        ages = np.ones(self._num_blocks)
        corrector_solves = 0
        enrichments = 0
        while True:
            U, eta, _, indicators = self.rd.sweep(mus, decompose=True)
            worst = int(np.argmax(eta))
            subs = self.reductor.d.solution_space.subspaces
            report(U, eta, worst, {'local_problem_solves': corrector_solves, 'global RB size': self.rd.solution_space.dim,
                                   'local RB sizes': [len(self.reductor.bases[s.id]) for s in subs]})
            if eta[worst] <= self.target_error:
                self.logger.info('estimated error %g is below the target %g', eta[worst], self.target_error)
                return U, eta
            if enrichments >= enrichment_steps:
                self.logger.warning('estimated error %g above target error of %g, but stopping since enrichment_steps=%s '
                                    'reached', eta[worst], self.target_error, enrichment_steps)
                return U, eta
            enrichments += 1
            try:
                corrector_solves = self._enrich(U[worst], mus[worst], indicators[:, worst], ages)
            except ExtensionError:
                if not stop_when_exhausted:
                    raise                                   # the reference lets it propagate (reductor.py:78)
                return U, eta                               # nothing new to add for the worst parameter

    def solve(self, mu, enrichment_steps=np.inf, callback=None):
        """reference ``online_enrichment.py:63-93``: ``(U, rd, reductor)`` for one parameter, ``callback(rd, U, mu, info)``
        after every solve with ``info = {'eta', 'local_problem_solves', 'global RB size', 'local RB sizes'}``.  An
        ``ExtensionError`` of a corrector that is already in the basis propagates, as it does there."""
        mu = self.discretization.parse_parameter(mu)

        def report(U, eta, worst, info):
            if callback:
                callback(self.rd, U[0], mu, dict(info, eta=float(eta[0])))
        U, _ = self._passes([mu], enrichment_steps, report, stop_when_exhausted=False)
        return U[0], self.rd, self.reductor

    def solve_batch(self, mus, enrichment_steps=np.inf, callback=None):
        """Enrichment driven by a whole parameter batch (no counterpart in the reference, which handles one ``mu`` per
        call): stops when the largest estimate over the batch is below ``target_error``, after ``enrichment_steps``
        enrichments, or when the worst parameter has nothing new to add.  ``callback(rd, U, mus, info)`` gets ``eta`` for
        the whole batch plus ``eta_max`` / ``argmax``.  Returns ``(U, eta, rd, reductor)`` for the final model, ``U`` / ``eta``
        covering the whole batch."""
        mus = [self.discretization.parse_parameter(mu) for mu in mus]

        def report(U, eta, worst, info):
            if callback:
                callback(self.rd, U, mus, dict(info, eta=eta, eta_max=float(eta[worst]), argmax=worst))
        U, eta = self._passes(mus, enrichment_steps, report, stop_when_exhausted=True)
        return U, eta, self.rd, self.reductor

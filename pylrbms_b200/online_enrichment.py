"""Online adaptive enrichment: Doerfler marking and the enrichment loop around the LRBMS hot path.

Drop-in for the reference's ``online_enrichment.py`` (``doerfler_marking`` ``:9-22``, ``AdaptiveEnrichment`` ``:25-93``):
same constructor arguments, same ``solve(mu, enrichment_steps, callback)`` return value ``(U, rd, reductor)``, same
marking rules (Doerfler on the *squared* indicators, plus every subdomain whose age exceeds ``marking_max_age``),
same age bookkeeping.  What it calls is the GPU path: ``rd.solve`` / ``rd.estimate`` (``lrbms_online_*``),
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
    """reference ``online_enrichment.py:25-93``."""

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

    def _enrich_once(self, U, mu, indicators, age_count):
        marked_subdomains = set(doerfler_marking(indicators, self.marking_doerfler_theta))
        num_dorfler_marked = len(marked_subdomains)
        self.logger.info('marked %d/%d subdomains due to Doerfler marking', num_dorfler_marked, self._num_blocks)
        for ii in np.where(age_count > self.marking_max_age)[0]:
            marked_subdomains.add(int(ii))
        self.logger.info('   and %d additionally due to age marking', len(marked_subdomains) - num_dorfler_marked)
        for ii in sorted(marked_subdomains):        # sorted: the reference iterates a set (order unspecified)
            self.reductor.enrich_local(ii, U, mu)
        self.rd = self.reductor.reduce()
        for ii in range(self._num_blocks):
            if ii in marked_subdomains:
                age_count[ii] = 1
            else:
                age_count[ii] += 1
        return len(marked_subdomains)

    def estimate(self, U, mu, decompose=False):
        return self.rd.estimate(U, mu=mu, decompose=decompose)

    def solve_batch(self, mus, enrichment_steps=np.inf, callback=None):
        """Enrichment driven by a whole parameter batch (no counterpart in the reference, which handles one ``mu`` per
        call, ``online_enrichment.py:63-93``): every pass sweeps ALL parameters in one ``rd.sweep`` (one launch set instead
        of ``len(mus)`` solve / estimate calls), takes the parameter with the largest estimate, marks subdomains from ITS
        indicators (Doerfler + age, the reference's rules), enriches with ITS solution and re-reduces; it stops when the
        largest estimate over the batch is below ``target_error``.  Returns ``(U, eta, rd, reductor)`` for the final model,
        ``U`` / ``eta`` covering the whole batch."""
        mus = [self.discretization.parse_parameter(mu) for mu in mus]
        age_count = np.ones(self._num_blocks)
        step, local_problem_solves = 1, 0
        from .reductor import ExtensionError
        while True:
            U, eta, _, indicators = self.rd.sweep(mus, decompose=True)
            worst = int(np.argmax(eta))
            if callback:
                subs = self.reductor.d.solution_space.subspaces
                callback(self.rd, U, mus, {'eta': eta, 'eta_max': float(eta[worst]), 'argmax': worst,
                                           'local_problem_solves': local_problem_solves,
                                           'global RB size': self.rd.solution_space.dim,
                                           'local RB sizes': [len(self.reductor.bases[s.id]) for s in subs]})
            if eta[worst] <= self.target_error or step > enrichment_steps:
                return U, eta, self.rd, self.reductor
            step += 1
            try:
                local_problem_solves = self._enrich_once(U[worst], mus[worst], indicators[:, worst], age_count)
            except ExtensionError:
                return U, eta, self.rd, self.reductor        # nothing new to add for the worst parameter

    def solve(self, mu, enrichment_steps=np.inf, callback=None):
        mu = self.discretization.parse_parameter(mu)
        enrichment_step = 1
        age_count = np.ones(self._num_blocks)
        local_problem_solves = 0
        rb_size = self.rd.solution_space.dim
        while True:
            U = self.rd.solve(mu)
            eta, _, indicators = self.estimate(U, mu=mu, decompose=True)
            indicators = np.asarray(indicators)
            if indicators.ndim == 2:
                indicators = indicators[:, 0]
            if callback:
                subs = self.reductor.d.solution_space.subspaces
                callback(self.rd, U, mu, {'eta': eta, 'local_problem_solves': local_problem_solves,
                                          'global RB size': self.rd.solution_space.dim,
                                          'local RB sizes': [len(self.reductor.bases[s.id]) for s in subs]})
            if eta <= self.target_error:
                self.logger.info('estimated error %g below target error of %g, no enrichment required', eta, self.target_error)
                return U, self.rd, self.reductor
            if enrichment_step > enrichment_steps:
                self.logger.warning('estimated error %g above target error of %g, but stopping since enrichment_steps=%s '
                                    'reached', eta, self.target_error, enrichment_steps)
                return U, self.rd, self.reductor
            enrichment_step += 1
            local_problem_solves = self._enrich_once(U, mu, indicators, age_count)
            self.logger.info('added %d local basis functions, system size increase: %d --> %d',
                             self.rd.solution_space.dim - rb_size, rb_size, self.rd.solution_space.dim)
            rb_size = self.rd.solution_space.dim

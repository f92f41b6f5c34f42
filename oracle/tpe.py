"""Oracle: TPE suggestion for `hp.uniform` spaces, restated from the published algorithm of hyperopt 0.2.4.

TEST INFRASTRUCTURE - see oracle/__init__.py.

PARITY UNPINNED.  The reference delegates all TPE arithmetic to the third-party package `hyperopt==0.2.4`
(requirements.txt:6; call sites optimisers/sequential/tpe_optimiser.py:51-53 and
hybrid_hyperband_tpe_with_transfer.py:116-158).  hyperopt is not vendored under /root/reference, is not installed in
this image and cannot be installed (no network), and no reference test pins anything at this boundary.  What follows
restates upstream hyperopt/tpe.py (`suggest`, `ap_filter_trials`, `adaptive_parzen_normal`,
`linear_forgetting_weights`, `GMM1`, `GMM1_lpdf`, `broadcast_best`), hyperopt/rand.py and the sequential
`fmin` loop (one suggestion per evaluation, a fresh seed per suggestion) from the published source as the author
remembers it; the only available pins are distributional: the reference's stored OFE vectors for the hybrids
(tests/golden/epdf_ofes.npz, `...sim(hb+tpe+transfer+*)...`).

Defaults of tpe.suggest: gamma = 0.25, n_EI_candidates = 24, prior_weight = 1.0, linear forgetting LF = 25; the
reference overrides n_startup_jobs = 10 + (number of injected trials) (tpe_optimiser.py:52).
"""
import math

import numpy as np

GAMMA, N_EI_CANDIDATES, PRIOR_WEIGHT, LF, EPS = 0.25, 24, 1.0, 25, 1e-12


def linear_forgetting_weights(n, lf=LF):
    """Ramp 1/n..1 over the oldest n - lf observations, 1 for the newest lf (tpe.linear_forgetting_weights)."""
    if n == 0:
        return np.asarray([])
    if n < lf:
        return np.ones(n)
    return np.concatenate([np.linspace(1.0 / n, 1.0, num=n - lf), np.ones(lf)])


def adaptive_parzen_normal(mus, prior_weight, prior_mu, prior_sigma, lf=LF):
    """Weights, sorted centres and bandwidths of the adaptive Parzen estimator (tpe.adaptive_parzen_normal):
    the prior is inserted among the sorted observations; each bandwidth is the larger gap to its neighbours, clipped to
    [prior_sigma / min(100, 1 + n_components), prior_sigma]; the prior keeps prior_sigma."""
    mus = np.asarray(mus, dtype=np.float64)
    n = len(mus)
    if n == 0:
        srtd, sigma, pos, order = np.asarray([prior_mu]), np.asarray([prior_sigma]), 0, np.zeros(0, int)
    elif n == 1:
        order = np.zeros(1, int)
        if prior_mu < mus[0]:
            pos, srtd, sigma = 0, np.asarray([prior_mu, mus[0]]), np.asarray([prior_sigma, prior_sigma * .5])
        else:
            pos, srtd, sigma = 1, np.asarray([mus[0], prior_mu]), np.asarray([prior_sigma * .5, prior_sigma])
    else:
        order = np.argsort(mus)
        pos = int(np.searchsorted(mus[order], prior_mu))
        srtd = np.zeros(n + 1)
        srtd[:pos], srtd[pos], srtd[pos + 1:] = mus[order[:pos]], prior_mu, mus[order[pos:]]
        sigma = np.zeros_like(srtd)
        sigma[1:-1] = np.maximum(srtd[1:-1] - srtd[0:-2], srtd[2:] - srtd[1:-1])
        sigma[0] = srtd[1] - srtd[0]
        sigma[-1] = srtd[-1] - srtd[-2]
    if lf and lf < n:
        unsrtd = linear_forgetting_weights(n, lf)
        w = np.zeros_like(srtd)
        w[:pos], w[pos], w[pos + 1:] = unsrtd[order[:pos]], prior_weight, unsrtd[order[pos:]]
    else:
        w = np.ones(len(srtd))
        w[pos] = prior_weight
    maxsigma = prior_sigma / 1.0
    minsigma = prior_sigma / min(100.0, (1.0 + len(srtd)))
    sigma = np.clip(sigma, minsigma, maxsigma)
    sigma[pos] = prior_sigma
    return w / w.sum(), srtd, sigma


def split_trials(vals, losses, gamma=GAMMA, gamma_cap=LF):
    """Observations of the n_below = min(ceil(gamma sqrt(N)), 25) lowest-loss trials ("good") and the rest, both in
    trial order (tpe.ap_filter_trials)."""
    losses = np.asarray(losses, dtype=np.float64)
    n_below = min(int(np.ceil(gamma * np.sqrt(len(losses)))), gamma_cap)
    order = np.argsort(losses, kind='stable')
    below = set(order[:n_below].tolist())
    vals = np.asarray(vals, dtype=np.float64)
    mask = np.asarray([i in below for i in range(len(losses))], dtype=bool)
    return vals[mask], vals[~mask]


def _normal_cdf(x, mu, sigma):
    z = (x - mu) / np.maximum(np.sqrt(2) * sigma, EPS)
    return 0.5 * (1 + np.vectorize(math.erf)(z))


def gmm1_sample(weights, mus, sigmas, low, high, size, rng):
    """Draws from the mixture truncated to [low, high) by rejection (tpe.GMM1)."""
    out = []
    while len(out) < size:
        active = int(np.argmax(rng.multinomial(1, weights)))
        draw = rng.normal(loc=mus[active], scale=sigmas[active])
        if low <= draw < high:
            out.append(draw)
    return np.asarray(out)


def gmm1_lpdf(samples, weights, mus, sigmas, low, high):
    """log-density of the truncated mixture (tpe.GMM1_lpdf)."""
    samples = np.asarray(samples, dtype=np.float64)
    p_accept = np.sum(weights * (_normal_cdf(high, mus, sigmas) - _normal_cdf(low, mus, sigmas)))
    dist = samples[:, None] - mus
    mahal = (dist / np.maximum(sigmas, EPS)) ** 2
    z = np.sqrt(2 * np.pi * sigmas ** 2)
    x = -0.5 * mahal + np.log(weights / z / p_accept)
    m = x.max(axis=1)
    return np.log(np.exp(x - m[:, None]).sum(axis=1)) + m


def suggest_uniform_dim(obs_vals, losses, low, high, rng, gamma=GAMMA, n_candidates=N_EI_CANDIDATES,
                        prior_weight=PRIOR_WEIGHT):
    """One hp.uniform(low, high) hyperparameter: candidates from the "good" model, best log l(x) - log g(x)
    (tpe.build_posterior + ap_uniform_sampler + broadcast_best).  Each hyperparameter is handled independently."""
    below, above = split_trials(obs_vals, losses, gamma)
    prior_mu, prior_sigma = 0.5 * (high + low), 1.0 * (high - low)
    wb, mb, sb = adaptive_parzen_normal(below, prior_weight, prior_mu, prior_sigma)
    wa, ma, sa = adaptive_parzen_normal(above, prior_weight, prior_mu, prior_sigma)
    cand = gmm1_sample(wb, mb, sb, low, high, n_candidates, rng)
    score = gmm1_lpdf(cand, wb, mb, sb, low, high) - gmm1_lpdf(cand, wa, ma, sa, low, high)
    return float(cand[int(np.argmax(score))])


class TpeRun:
    """The sequential fmin loop of tpe_optimiser.py:51-53 for a space of independent hp.uniform hyperparameters:
    max_queue_len = 1, so suggestion k sees the losses of trials 0..k-1; a new RandomState per suggestion seeded
    from the run's own RandomState (fmin.FMinIter.run); random suggestions until 10 + n_injected trials exist."""

    def __init__(self, bounds, injected=(), rstate=None, n_startup_base=10):
        self.bounds = list(bounds)                             # [(low, high)] per hyperparameter
        self.vals = [list(v) for v, _ in injected]             # injected trials first (they carry the lowest tids)
        self.losses = [float(l) for _, l in injected]
        self.n_injected = len(self.losses)
        self.n_startup = n_startup_base + self.n_injected
        self.rstate = rstate if rstate is not None else np.random.RandomState()

    def suggest(self):
        rng = np.random.RandomState(self.rstate.randint(2 ** 31 - 1))
        if len(self.losses) < self.n_startup:                  # rand.suggest
            return [float(rng.uniform(lo, hi)) for lo, hi in self.bounds]
        obs = np.asarray(self.vals, dtype=np.float64)
        return [suggest_uniform_dim(obs[:, d], self.losses, lo, hi, rng) for d, (lo, hi) in enumerate(self.bounds)]

    def observe(self, vals, loss):
        self.vals.append(list(vals))
        self.losses.append(float(loss))

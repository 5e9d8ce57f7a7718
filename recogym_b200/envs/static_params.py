"""Static parameters of reco-gym-v1, generated on the HOST with NumPy's legacy RandomState so that
Gamma, mu_organic, beta, mu_bandit and the transition matrix are bit-identical to the reference
for the same `random_seed` (one-time work).

Follows RecoEnv1.set_static_params / generate_beta / normalize_beta
(reference: recogym/envs/reco_env_v1.py:51-75,130-174).
"""
import numpy as np
from numpy.random.mtrand import RandomState

ORGANIC, BANDIT, STOP = 0, 1, 2


def transition_matrix(config):
    """3x3 organic/bandit/stop chain (reco_env_v1.py:54-61).  NB the bandit row leaves with
    prob_leave_ORGANIC -- prob_leave_bandit is unused by the reference, so it is unused here."""
    m = np.array([
        [0, config.prob_organic_to_bandit, config.prob_leave_organic],
        [config.prob_bandit_to_organic, 0, config.prob_leave_organic],
        [0.0, 0.0, 1.0],
    ])
    m[0, 0] = 1 - sum(m[0, :])
    m[1, 1] = 1 - sum(m[1, :])
    return m


def flip_index(Gamma, number_of_flips):
    """Row permutation that swaps the `number_of_flips` most correlated product pairs, each product
    used at most once (reco_env_v1.py:142-165).  O(P^2 log P): host-only, small P."""
    P = Gamma.shape[0]
    index = np.arange(P)
    cov = Gamma @ Gamma.T
    cov = cov - np.diag(np.diag(cov))
    used = set()
    done = 0
    for flat in cov.flatten().argsort()[::-1]:
        i, j = int(flat / P), np.mod(flat, P)
        if i in used or j in used:
            continue
        index[i], index[j] = j, i
        used.update((i, j))
        done += 1
        if done == number_of_flips:
            break
    return index


BLOCKED_FLIPS_MIN_PRODUCTS = 8192   # above this the P x P covariance (8 P^2 bytes) and its argsort give way to row maxima


def flip_index_blocked(Gamma, number_of_flips, block=2048):
    """flip_index for catalogues whose P x P covariance does not fit (P = 100 000: 80 GB + an argsort of 10^10 entries;
    SURVEY.md 8f rank 4).  The reference walks ALL pairs in descending covariance and flips a pair when neither product
    has been used (reco_env_v1.py:152-165): the pair it flips next is always the maximum of the covariance over the
    still-unused pairs, i.e. the largest of the unused rows' maxima over unused columns.  So: one blocked pass for every
    row's best partner (O(P^2 K) flops, O(P block) memory), then per flip only the rows whose best partner has just been
    used are recomputed.  Identical to flip_index unless two DIFFERENT pairs tie to the last bit (the symmetric twins
    (i, j) / (j, i) select the same flip whichever comes first)."""
    P = Gamma.shape[0]
    index = np.arange(P)
    unused = np.ones(P, bool)
    best_val = np.full(P, -np.inf)
    best_j = np.full(P, -1, np.int64)

    def row_best(rows):
        cov = Gamma[rows] @ Gamma.T                           # [len(rows), P], the rows of the reference's prod_cov
        cov[np.arange(len(rows)), rows] = 0.0                 # prod_cov - diag(diag(prod_cov)): a zero, not -inf
        cov[:, ~unused] = -np.inf
        cov[~unused[rows]] = -np.inf
        # the reference's descending walk meets equal values in the REVERSE of argsort's order; exact ties between different
        # pairs are the caveat above, any maximiser will do
        j = cov.argmax(axis=1)
        best_val[rows] = cov[np.arange(len(rows)), j]
        best_j[rows] = j

    for b0 in range(0, P, block):
        row_best(np.arange(b0, min(P, b0 + block)))
    for _ in range(int(number_of_flips)):
        i = int(np.argmax(np.where(unused, best_val, -np.inf)))
        if not np.isfinite(best_val[i]):
            break                                              # fewer than two unused products left
        j = int(best_j[i])
        index[i], index[j] = j, i
        unused[i] = unused[j] = False
        best_val[i] = best_val[j] = -np.inf
        stale = np.flatnonzero(unused & ((best_j == i) | (best_j == j)))
        for b0 in range(0, len(stale), block):
            row_best(stale[b0:b0 + block])
    return index


def make_beta(Gamma, mu_organic, number_of_flips, normalize):
    """beta / mu_bandit from Gamma / mu_organic (reco_env_v1.py:133-174).  With zero flips beta IS
    Gamma (same array object in the reference, :137-138)."""
    if number_of_flips == 0:
        beta, mu_bandit = Gamma, mu_organic
    else:
        idx = (flip_index_blocked if Gamma.shape[0] >= BLOCKED_FLIPS_MIN_PRODUCTS else flip_index)(Gamma, number_of_flips)
        beta, mu_bandit = Gamma[idx, :], mu_organic[idx, :]
    if normalize:
        beta = beta / np.sqrt((beta ** 2).sum(1)[:, np.newaxis])  # reco_env_v1.py:130-131
    return beta, mu_bandit


def draw_static_params(config, rng=None):
    """Returns dict(state_transition, Gamma, mu_organic, beta, mu_bandit) drawn from
    RandomState(config.random_seed) in the reference's order (Gamma first, then mu_organic)."""
    if rng is None:
        rng = RandomState(config.random_seed)
    P, K = config.num_products, config.K
    Gamma = rng.normal(size=(P, K))
    mu_organic = rng.normal(0, config.sigma_mu_organic, size=(P, 1))
    beta, mu_bandit = make_beta(Gamma, mu_organic, config.number_of_flips, config.normalize_beta)
    return {
        "state_transition": transition_matrix(config),
        "Gamma": Gamma, "mu_organic": mu_organic, "beta": beta, "mu_bandit": mu_bandit,
    }

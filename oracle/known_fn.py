"""Oracle: closest-known-loss-function approximation (reference autotune/benchmarks/known_loss_fn_problem.py,
autotune/core/arm.py:57-70, autotune/core/params.py:30-47).

TEST INFRASTRUCTURE - see oracle/__init__.py.  A problem's domain is a list of dicts
  {name, min_val, max_val, scale ('log'|'linear'), interval (None|number), init_val, optimised (bool)}
in domain order; an arm is a dict name -> value whose key order is the arm's attribute order in the reference
(non-optimised defaults first, then the optimised hyperparameters, each in domain order - arm.py:29-55).
"""
import numpy as np

from . import hyperband as ohb

MNIST_DOMAIN = [  # benchmarks/mnist_problem.py:11-19
    dict(name='learning_rate', min_val=np.log(10 ** -6), max_val=np.log(10 ** 0), scale='log', interval=None, init_val=None, optimised=True),
    dict(name='weight_decay', min_val=np.log(10 ** -6), max_val=np.log(10 ** -1), scale='log', interval=None, init_val=None, optimised=True),
    dict(name='momentum', min_val=0.3, max_val=0.999, scale='linear', interval=None, init_val=None, optimised=True),
    dict(name='batch_size', min_val=20, max_val=2000, scale='linear', interval=1, init_val=None, optimised=True),
]
_U = dict(min_val=np.log(2 ** 4), max_val=np.log(2 ** 8), scale='log', interval=1, init_val=None, optimised=True)
CIFAR_DOMAIN = [  # benchmarks/cifar_problem.py:11-33 (5 of 9 hyperparameters optimised)
    dict(name='learning_rate', min_val=np.log(10 ** -6), max_val=np.log(10 ** 0), scale='log', interval=None, init_val=None, optimised=True),
    dict(name='n_units_1', **_U), dict(name='n_units_2', **_U), dict(name='n_units_3', **_U),
    dict(name='batch_size', min_val=32, max_val=512, scale='linear', interval=1, init_val=None, optimised=True),
    dict(name='lr_step', min_val=1, max_val=5, scale='linear', interval=1, init_val=1, optimised=False),
    dict(name='gamma', min_val=np.log(10 ** -3), max_val=np.log(10 ** -1), scale='log', interval=None, init_val=0.1, optimised=False),
    dict(name='weight_decay', min_val=np.log(10 ** -6), max_val=np.log(10 ** -1), scale='log', interval=None, init_val=0.004, optimised=False),
    dict(name='momentum', min_val=0.3, max_val=0.999, scale='linear', interval=None, init_val=0.9, optimised=False),
]


def attribute_order(domain):
    """Key order of an arm made by Arm().draw_hp_val (arm.py:29-55): defaults first, then the drawn ones."""
    return [p['name'] for p in domain if not p['optimised']] + [p['name'] for p in domain if p['optimised']]


def draw_arm(domain):
    """Arm.draw_hp_val with numpy's global RNG (arm.py:44-55, params.py:36-45)."""
    arm = {p['name']: p['init_val'] for p in domain if not p['optimised']}
    for p in domain:
        if not p['optimised']:
            continue
        val = np.random.rand(1) * (p['max_val'] - p['min_val']) + p['min_val']
        if p['scale'] == 'log':
            val = np.array([np.e ** v for v in val])
        if p['interval']:
            val = (np.floor(val / p['interval']) * p['interval']).astype(int)
        arm[p['name']] = val[0]
    return arm


def normalise(arm, domain):
    """Arm.normalize (arm.py:57-70): (v - min_val) / (max_val - min_val); log-scale bounds are exponents while v is
    linear - the reference's quirk, kept."""
    by_name = {p['name']: p for p in domain}
    return {n: (arm[n] - by_name[n]['min_val']) / (by_name[n]['max_val'] - by_name[n]['min_val'])
            for n in arm if n in by_name}


def nearest(arm, db_arms, domain):
    """Index of the recorded arm with the smallest normalised squared error (known_loss_fn_problem.py:35-50):
    python sum() over the proposed arm's attributes in order, strict `<` so the first minimum wins."""
    q = normalise(arm, domain)
    best, best_i = np.inf, None
    for i, rec in enumerate(db_arms):
        r = normalise(rec, domain)
        err = sum((float(q[n]) - float(r[n])) ** 2 for n in arm)
        if err < best:
            best, best_i = err, i
    return best_i, best


def evaluate(arm, db_arms, curves, domain, n_resources):
    """KnownFnEvaluator.evaluate (known_loss_fn_problem.py:22-62): loss of the nearest recorded curve at time-1."""
    i, _ = nearest(arm, db_arms, domain)
    return curves[i][int(n_resources) - 1]


def run_hyperband(db_arms, curves, domain, R, eta, arms=None, min_or_max='min'):
    """One Hyperband repetition on the closest-known-function problem (run_closest_loss_fn_approximation.py:183-194).
    `arms` (list of dicts in creation order) replaces the global-RNG draws when given.  Returns a dict like
    oracle.hyperband.run_predrawn (survivor ids, round-bests, OFE, nearest index per arm)."""
    plan = ohb.bracket_plan(R, eta)
    desc = min_or_max == 'max'
    created, nn_idx, rb, rb_id, surv_ids, eval_ids = [], [], [], [], [], []
    nxt = 0
    for b in plan:
        alive = []
        for _ in range(b['n']):
            arm = arms[nxt] if arms is not None else draw_arm(domain)
            created.append(arm)
            nn_idx.append(nearest(arm, db_arms, domain)[0])
            alive.append(nxt)
            nxt += 1
        for rd in b['rounds']:
            losses = [curves[nn_idx[a]][int(rd['resource']) - 1] for a in alive]
            surv, best_id, best = ohb.halve(alive, losses, rd['keep'], desc)
            eval_ids += alive
            surv_ids += surv
            rb.append(best)
            rb_id.append(best_id)
            alive = surv
    pick = max if desc else min
    j = pick(range(len(rb)), key=lambda i: rb[i])
    return dict(arms=created, nn_idx=np.asarray(nn_idx, np.int32), surv_ids=np.asarray(surv_ids, np.int32),
                round_best=np.asarray(rb, np.float64), round_best_id=np.asarray(rb_id, np.int32),
                ofe=np.float64(rb[j]), ofe_id=np.int32(rb_id[j]))


def synthetic_db(domain, n_arms, curve_len, seed):
    """Synthetic recorded-arm database of SURVEY 8d config 4: arms drawn like the reference draws them (numpy global
    RNG seeded with `seed`), curves c_a(t) = b_a + (1 - b_a) exp(-t / tau_a), b ~ U[0.07, 0.57], tau ~ U[5, 55]
    from default_rng(seed)."""
    state = np.random.get_state()
    np.random.seed(seed)
    db_arms = [draw_arm(domain) for _ in range(n_arms)]
    np.random.set_state(state)
    rng = np.random.default_rng(seed)
    b = rng.uniform(0.07, 0.57, n_arms)
    tau = rng.uniform(5, 55, n_arms)
    t = np.arange(1, curve_len + 1)
    curves = b[:, None] + (1 - b[:, None]) * np.exp(-t[None, :] / tau[:, None])
    return db_arms, curves

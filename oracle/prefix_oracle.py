"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's legacy prefix search (decoding/prefix_search.py and
the Cython helpers it calls in decoding_cy.pyx), used by tests/ to check csrc/prefix.cu.  Pinned against the
unmodified reference package run in this container: tests/golden/prefix_golden.npz (tests/golden/make_prefix_golden.py).

Two arithmetic flavours, as in the reference: "numpy" (np.logaddexp, scipy-style logsumexp, LOG_0 = -inf) and "cy"
(log(exp(a) + exp(b)), forward vectors initialised to -9999: decoding_cy.pyx:18, :127-156, :177-220)."""
import numpy as np

LOG_0 = -float("inf")
LOG_1 = 0.0


def _logsumexp(a):
    """scipy.special.logsumexp on a flat array: shift by the maximum (0 when it is not finite)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    if a.size == 0:
        return LOG_0
    m = np.max(a)
    if not np.isfinite(m):
        m = 0.0
    with np.errstate(divide="ignore"):
        return float(np.log(np.sum(np.exp(a - m))) + m)


def _lae(flavour, x, y):
    if flavour == "cy":
        with np.errstate(divide="ignore"):
            return float(np.log(np.exp(x) + np.exp(y)))  # decoding_cy.pyx:154
    return float(np.logaddexp(x, y))


def forward_vec_log(s, i, y, previous=None, flavour="numpy"):
    """prefix_search.py:81-97 / decoding_cy.pyx:127-156"""
    T = len(y)
    fw = np.zeros(T) + (-9999.0 if flavour == "cy" else LOG_0)
    for t in range(T):
        if i == 0:
            fw[t] = y[t, s] if t == 0 else y[t, -1] + fw[t - 1]
        elif t == 0:
            if i == 1:
                fw[t] = y[t, s]
        else:
            fw[t] = _lae(flavour, y[t, -1] + fw[t - 1], y[t, s] + previous[t - 1])
    return fw


def forward_vec_no_gap_log(l, y, fw0):
    """prefix_search.py:67-79"""
    return np.insert(fw0[:-1], 0, LOG_1 if len(l) == 1 else LOG_0) + y[:, l[-1]]


def prefix_search(y, n_letters, flavour="numpy"):
    """prefix_search.py:116-174 / :176-238: (label indices, log label probability)."""
    y = np.asarray(y, dtype=np.float64)
    top, cur = (), ()
    label_prob = {(): float(np.sum(y[:, -1]))}
    alpha_prev = forward_vec_log(-1, 0, y, flavour=flavour)
    level = 0
    while True:
        level += 1
        prefix_prob, alphas = {}, []
        best = None
        for c in range(n_letters):
            prefix = cur + (c,)
            if c == 0:
                best = prefix
            prefix_prob[prefix] = _logsumexp(forward_vec_no_gap_log(prefix, y, alpha_prev))
            alpha = forward_vec_log(c, level, y, previous=alpha_prev, flavour=flavour)
            label_prob[prefix] = alpha[-1]
            if label_prob[prefix] > label_prob[top]:
                top = prefix
            if prefix_prob[prefix] > prefix_prob[best]:
                best = prefix
            alphas.append(alpha)
        if prefix_prob[best] < label_prob[top] or level >= len(y) + 2:
            break
        cur = best
        alpha_prev = alphas[cur[-1]]
    return list(top), float(label_prob[top])


def pair_gamma(y1, y2, flavour="numpy"):
    """prefix_search.py:35-65 / decoding_cy.pyx:177-220"""
    y1 = np.asarray(y1, dtype=np.float64)
    y2 = np.asarray(y2, dtype=np.float64)
    U, V = len(y1), len(y2)
    L0 = -9999.0 if flavour == "cy" else LOG_0
    g = np.zeros((U + 1, V + 1)) + L0
    ga = np.zeros((U + 1, V + 1)) + L0
    g[U, V] = LOG_1
    ga[U, V] = LOG_1
    for v in range(V):
        s = 0.0
        for k in range(v, V):
            s += y2[k, -1]
        g[U, v] = s
    for u in range(U):
        s = 0.0
        for k in range(u, U):
            s += y1[k, -1]
        g[u, V] = s
    for u in reversed(range(U)):
        for v in reversed(range(V)):
            gamma_eps = g[u + 1, v] + y1[u, -1]
            gamma_ast_eps = ga[u, v + 1] + y2[v, -1]
            if flavour == "cy":
                with np.errstate(divide="ignore"):
                    tot = float(np.log(np.sum(np.exp(y1[u, :-1] + y2[v, :-1]))))
            else:
                tot = _logsumexp(y1[u, :-1] + y2[v, :-1])
            gamma_ast_ast = g[u + 1, v + 1] + tot
            ga[u, v] = _lae(flavour, gamma_ast_eps, gamma_ast_ast)
            g[u, v] = _lae(flavour, gamma_eps, ga[u, v])
    return g


def pair_prefix_search(y1, y2, n_letters, flavour="numpy"):
    """prefix_search.py:247-310 / :312-385: (label indices, log label probability)."""
    y1 = np.asarray(y1, dtype=np.float64)
    y2 = np.asarray(y2, dtype=np.float64)
    gamma = pair_gamma(y1, y2, flavour)
    stop = False
    level = 0
    top, cur = (), ()
    label_prob = {(): float(np.sum(y1[:, -1]) + np.sum(y2[:, -1]))}
    a1_prev = forward_vec_log(-1, 0, y1, flavour=flavour)
    a2_prev = forward_vec_log(-1, 0, y2, flavour=flavour)
    while not stop:
        prefix_prob, alphas = {}, []
        level += 1
        if len(cur) > max(len(y1), len(y2)):
            stop = True
        for c in range(n_letters):
            prefix = cur + (c,)
            ast1 = forward_vec_no_gap_log(prefix, y1, a1_prev)
            ast2 = forward_vec_no_gap_log(prefix, y2, a2_prev)
            prefix_prob[prefix] = _logsumexp((np.add.outer(ast1, ast2) + gamma[1:, 1:]).flatten()) - gamma[0, 0]
            a1 = forward_vec_log(c, level, y1, previous=a1_prev, flavour=flavour)
            a2 = forward_vec_log(c, level, y2, previous=a2_prev, flavour=flavour)
            label_prob[prefix] = a1[-1] + a2[-1] - gamma[0, 0]
            alphas.append((a1, a2))
        best = max(prefix_prob.items(), key=lambda kv: kv[1])[0]
        if prefix_prob[best] < label_prob[top]:
            stop = True
        else:
            top = max(label_prob.items(), key=lambda kv: kv[1])[0]
            cur = best
            a1_prev, a2_prev = alphas[cur[-1]]
    return list(top), float(label_prob[top])

"""TEST INFRASTRUCTURE ONLY: numpy restatement of the NMF solvers the reference's `regularized_nmf` (utilities.py:253-299) delegates
to, for the initialisation-pipeline row.

The algorithm lives in a third-party dependency that is not part of /root/reference: scikit-learn (`sklearn.decomposition.NMF`,
version 1.9.0 in this image; the reference's pyproject pins none).  What is restated: `_fit_multiplicative_update` for
beta_loss in {'frobenius', 'kullback-leibler'} (Fevotte & Idier 2011 MM updates with sklearn's EPSILON guards and its
every-10-iterations stopping rule) and `_fit_coordinate_descent` (Cichocki & Phan 2009 / Hsieh & Dhillon 2011 cyclic coordinate
descent, components in order, projected-gradient stopping rule).  Pinned against sklearn itself in tests/test_oracle.py; the CUDA
path (gpzoo_b200/initialisation.py) is compared with sklearn directly and with the reference's outputs in tests/test_init_gpu.py."""
import numpy as np

EPSILON = np.finfo(np.float32).eps


def beta_divergence(X, W, H, beta):
    if beta == 2:
        return np.sqrt(((X - W @ H) ** 2).sum())
    WH = (W @ H).ravel()
    Xr = X.ravel()
    nz = Xr > EPSILON
    WHn = np.maximum(WH[nz], EPSILON)
    res = Xr[nz] @ np.log(Xr[nz] / WHn) + W.sum(0) @ H.sum(1) - Xr[nz].sum()
    return np.sqrt(2 * max(res, 0.0))


def nmf_mu(X, W, H, beta, max_iter=200, tol=1e-4):
    """-> (W, H, n_iter); beta = 2 (Frobenius) or 1 (Kullback-Leibler)."""
    W, H = W.copy(), H.copy()
    err0 = prev = beta_divergence(X, W, H, beta)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        if beta == 2:
            num, den = X @ H.T, W @ (H @ H.T)
        else:
            num, den = (X / np.maximum(W @ H, EPSILON)) @ H.T, np.broadcast_to(H.sum(1)[None, :], W.shape).copy()
        den[den == 0] = EPSILON
        W *= num / den
        if beta == 2:
            num, den = W.T @ X, W.T @ W @ H
        else:
            ws = W.sum(0)
            ws[ws == 0] = 1.0
            num, den = W.T @ (X / np.maximum(W @ H, EPSILON)), np.broadcast_to(ws[:, None], H.shape).copy()
        den[den == 0] = EPSILON
        H *= num / den
        if beta <= 1:
            H[H < np.finfo(np.float64).eps] = 0.0
        if tol > 0 and n_iter % 10 == 0:
            err = beta_divergence(X, W, H, beta)
            if (prev - err) / err0 < tol:
                break
            prev = err
    return W, H, n_iter


def _cd_sweep(W, HHt, XHt):
    violation = 0.0
    for t in range(W.shape[1]):
        grad = W @ HHt[:, t] - XHt[:, t]
        pg = np.where(W[:, t] == 0, np.minimum(grad, 0), grad)
        violation += np.abs(pg).sum()
        if HHt[t, t] != 0:
            W[:, t] = np.maximum(W[:, t] - grad / HHt[t, t], 0)
    return violation


def nmf_cd(X, W, H, max_iter=200, tol=1e-4):
    W, Ht = W.copy(), H.T.copy()
    v0, n_iter = None, 0
    for n_iter in range(1, max_iter + 1):
        v = _cd_sweep(W, Ht.T @ Ht, X @ Ht) + _cd_sweep(Ht, W.T @ W, X.T @ W)
        if n_iter == 1:
            v0 = v
        if v0 == 0 or v / v0 <= tol:
            break
    return W, Ht.T, n_iter

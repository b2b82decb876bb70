"""Design study (CPU only): how many bits of a sequence weight does one u8 x s8 tensor-core pass carry?

The contraction represents a weight as wq = u * V with an 8-bit multiplier u on the one-hot operand and S base-256 digits of V on
the weighted operand (one pass per digit): ~8 S + 5 bits.  If every pass had its OWN pair (u_k, V_k), wq = sum_k m_k u_k V_k with
integer gains m_k (residual passes signed), each pass would approximate the residual of the previous ones by the best of 255
products -- this script measures what that buys on the bench weights.  Result on the SSU bench weights: 3 such passes reach the
precision of S = 4 (37.8 vs 37.3 bits below the largest weight), 2 passes 25-26 bits (S = 2: 21.3, S = 3: 28.9).  The price is a
different one-hot operand per pass, i.e. N = 4 CJ per MMA instead of 4 S CJ for the same A tile ("What comes next")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

U = np.arange(1, 256).astype(np.float64)


def candidates(T, vmin, vmax, topk):
    """residuals T - u V of the top-k products per element, u in 1..255, V in [vmin, vmax]"""
    V = np.clip(np.rint(T[:, None] / U[None, :]), vmin, vmax)
    R = T[:, None] - U[None, :] * V
    k = np.argsort(np.abs(R), axis=1)[:, :topk]
    return R[np.arange(len(T))[:, None], k]


def independent_products(w, passes, f1, beam):
    T = w / w.max() * (255 * 255 * f1)                 # pass-1 targets: the largest weight at a fraction f1 of the product range
    R = candidates(T, 0, 255, beam)
    unit = 1.0
    for p in range(1, passes):
        m = np.abs(R).min(1).max()
        if m == 0:
            break
        g = max(np.floor(255 * 127 / m), 1.0)          # integer gain of the next (signed) pass
        R = np.concatenate([candidates(R[:, b] * g, -127, 127, beam if p < passes - 1 else 1) for b in range(R.shape[1])], axis=1)
        k = np.argsort(np.abs(R), axis=1)[:, :beam]
        R = R[np.arange(len(T))[:, None], k]
        unit *= g
    err = (np.abs(R).min(1) / unit).max() / (255 * 255 * f1) * w.max()
    return np.log2(w.max() / err)


def shared_multiplier(w, S):
    vlim = 256.0 ** S
    q = np.floor(np.log2(255 * (vlim - 1) / w.max()))
    T = (w * 2.0 ** q)[:, None]
    V = np.rint(T / U[None, :])
    err = np.where(V < vlim, np.abs(T - U[None, :] * V), np.inf).min(1)
    return np.log2(w.max() / (err * 2.0 ** -q).max())


if __name__ == "__main__":
    synth = ge.load_package().synth
    _, wgt, _, _ = synth.synthetic_family(10000, 1800, seed=42)
    for S in (1, 2, 3, 4):
        print(f"shared multiplier (current), S = {S}: {shared_multiplier(wgt, S):.1f} bits below the largest weight")
    for passes in (2, 3):
        best = max((independent_products(wgt, passes, f1, 4), f1) for f1 in (1.0, 0.7, 0.5, 0.35))
        print(f"independent products, {passes} passes: {best[0]:.1f} bits (largest weight at {best[1]} of the product range)")

"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's token sampling (egom2p/models/generate.py:332-371):
`top_k_top_p_filtering` (top-k, then nucleus filtering on the softmax at temperature 1 with the "shift right" rule that keeps
the first token crossing top_p) followed by softmax(kept / temperature). torch.multinomial's random stream cannot be
reproduced outside torch, so the draw itself is restated as an inverse-CDF walk in ascending token order over a supplied
uniform -- the definition include/egom2p_b200.h gives for egom2p_sample_rows. Pinned against the live reference's filter by
tests/golden/sampling_filter.npz (oracle/gen_golden_sampling.py, tests/test_oracle.py)."""
import numpy as np


def kept_mask(logits: np.ndarray, top_k=0.0, top_p=0.0) -> np.ndarray:
    """(rows, V) float32 -> bool mask of the tokens that survive top-k / top-p (generate.py:332-359)."""
    logits = np.asarray(logits, dtype=np.float32)
    rows, V = logits.shape
    keep = np.ones((rows, V), dtype=bool)
    work = logits.astype(np.float64).copy()
    if top_k > 0.0:
        k = min(top_k, V) if isinstance(top_k, int) else min(int(top_k * V), V)
        kth = np.sort(logits, axis=1)[:, ::-1][:, k - 1][:, None]        # torch.topk(logits, k)[0][..., -1, None]
        keep &= ~(logits < kth)
        work[~keep] = -np.inf
    if top_p > 0.0:
        order = np.argsort(-work, axis=1, kind="stable")
        srt = np.take_along_axis(work, order, axis=1)
        e = np.exp(srt - srt[:, :1])
        cum = np.cumsum(e / e.sum(1, keepdims=True), axis=1)
        remove = cum > top_p
        remove[:, 1:] = remove[:, :-1].copy()                            # keep the first token above the threshold
        remove[:, 0] = False
        rm = np.zeros_like(remove)
        np.put_along_axis(rm, order, remove, axis=1)
        keep &= ~rm
    return keep


def probs(logits: np.ndarray, temperature: float, top_k=0.0, top_p=0.0) -> np.ndarray:
    keep = kept_mask(logits, top_k, top_p)
    z = np.where(keep, logits.astype(np.float64) / temperature, -np.inf)
    z -= z.max(1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(1, keepdims=True)


def draw(logits: np.ndarray, temperature: float, u: np.ndarray, top_k=0.0, top_p=0.0):
    """Inverse-CDF draw in ascending token order. Returns (token, prob, margin): margin = distance of u from the nearest CDF
    step, so that tests can skip draws that sit on a rounding boundary."""
    if abs(temperature) <= 1e-10:
        tok = logits.argmax(1)
        return tok, np.ones(len(tok)), np.ones(len(tok))
    p = probs(logits, temperature, top_k, top_p)
    cdf = np.cumsum(p, axis=1)
    tok = np.array([min(int(np.searchsorted(cdf[r], u[r], side="right")), p.shape[1] - 1) for r in range(len(u))])
    # a draw that lands beyond the last kept token (rounding) goes to the last kept token
    for r in range(len(u)):
        if p[r, tok[r]] == 0.0:
            tok[r] = np.nonzero(p[r])[0][-1]
    lo = np.where(tok > 0, cdf[np.arange(len(u)), np.maximum(tok - 1, 0)], 0.0)
    hi = cdf[np.arange(len(u)), tok]
    margin = np.minimum(u - lo, hi - u)
    return tok, p[np.arange(len(u)), tok], margin

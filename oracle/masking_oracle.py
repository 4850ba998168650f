"""TEST INFRASTRUCTURE ONLY. CPU restatement of UnifiedMasking.image_mask (egom2p/data/masking.py:236-266) as a function of
the noise it draws. The reference gathers a [0]*budget + [1]*rest vector THROUGH ids_shuffle = argsort(noise), i.e.
input_mask[i] = (ids_shuffle[i] >= input_budget): position i is an input iff the element of rank i has an index below the
budget (the inverse permutation of the usual "first k of a shuffle"; just as uniform). decoder_attention_mask = target
count at the lowest target position. Pinned by tests/golden/masking_ref.npz."""
import numpy as np


def image_mask(noise: np.ndarray, input_budget: int, target_budget):
    L = noise.shape[0]
    order = np.argsort(noise, kind="stable")
    input_mask = order >= input_budget
    if target_budget is None:
        target_mask = ~input_mask
    else:
        target_mask = ~((order >= input_budget) & (order < input_budget + target_budget))
    attn = np.zeros(L, dtype=np.int32)
    first = int(np.argmin(target_mask.astype(np.float32) + np.arange(L, dtype=np.float32) * np.float32(1e-6)))
    attn[first] = int((~target_mask).sum())
    return input_mask, target_mask, attn

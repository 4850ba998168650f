"""The four pre-tokenised modalities of ego-b "mod4" (reference: egom2p/data/modality_info.py:59-69,75-85,116-141).
ids are sha256(name) % 2**15 (egom2p/utils/misc.py:39-41) -- recomputed here and checked in tests."""
import hashlib
from functools import partial

from .adapters import (GazeCamTokenDecoderEmbedding, GazeCamTokenEncoderEmbedding, VideoTokenDecoderEmbedding,
                       VideoTokenEncoderEmbedding)


def generate_uint15_hash(seed_str: str) -> int:
    return int(hashlib.sha256(seed_str.encode("utf-8")).hexdigest(), 16) % (2 ** 15)


def _video(name, path):
    return {"input_size": 256, "patch_size": 8, "vocab_size": 64000, "min_tokens": 0, "max_tokens": 5120, "type": "img",
            "id": generate_uint15_hash(name), "pretokenized": True, "path": path,
            "encoder_embedding": partial(VideoTokenEncoderEmbedding, vocab_size=64000),
            "decoder_embedding": partial(VideoTokenDecoderEmbedding, vocab_size=64000)}


def _small(name, kind):
    return {"vocab_size": 256, "min_tokens": 0, "max_tokens": 30, "type": kind, "id": generate_uint15_hash(name),
            "pretokenized": True, "path": kind,
            "encoder_embedding": partial(GazeCamTokenEncoderEmbedding, vocab_size=256),
            "decoder_embedding": partial(GazeCamTokenDecoderEmbedding, vocab_size=256)}


MODALITY_INFO = {
    "tok_rgb": _video("tok_rgb", "rgb"),
    "tok_depth": _video("tok_depth", "depth"),
    "tok_cam": _small("tok_cam", "cam"),
    "tok_gaze": _small("tok_gaze", "gaze"),
}
MOD4 = sorted(MODALITY_INFO)  # dict order used by run_training_egom2p.py:274-276: cam, depth, gaze, rgb

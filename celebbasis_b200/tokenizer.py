"""Stand-in for transformers.CLIPTokenizer when the CLIP BPE vocabulary is not on disk (no network here).

The real tokenizer is third-party host code (transformers==4.18.0, environment.yaml:199) called at
ldm/modules/encoders/modules.py:460-462; FrozenCLIPEmbedder in the host mirror uses it whenever
`CLIPTokenizer.from_pretrained` succeeds.  This deterministic whitespace tokenizer keeps the properties the
path relies on: (1, 77) int64 ids, BOS 49406, EOS/PAD 49407, one id per placeholder word
(embedding_manager.py:18-24 asserts that), and the real ids the reference hard-codes for its special words
(modules.py:198-231,259-262; 02_start_test.sh:106-107).
"""
import hashlib

import torch

KNOWN_TOKENS = {
    "sks": 48136, "ks": 662, "ata": 4236, "tre": 6033, "a": 320, "photo": 1125, "of": 539,
    "elon": 20406, "musk": 19063,
}
BOS, EOS = 49406, 49407


class SyntheticCLIPTokenizer:
    def __init__(self, max_length=77, phrases=None):
        """phrases: optional {text: [ids]} table of REAL CLIP BPE ids for whole strings (the celebrity names of
        infer_images/token_len.txt), looked up before the per-word fallback."""
        self.model_max_length = max_length
        self.phrases = dict(phrases or {})

    @staticmethod
    def word_id(word):
        w = word.lower()
        if w in KNOWN_TOKENS:
            return KNOWN_TOKENS[w]
        h = int.from_bytes(hashlib.sha1(w.encode()).digest()[:4], "little")
        return 1000 + h % 40000

    def encode_one(self, text, max_length):
        body = self.phrases.get(text)
        if body is None:
            body = [self.word_id(w) for w in text.split()]
        ids = [BOS] + list(body)[: max_length - 2] + [EOS]
        return ids + [EOS] * (max_length - len(ids))

    def __call__(self, text, truncation=True, max_length=77, return_length=True, return_overflowing_tokens=False,
                 padding="max_length", return_tensors="pt"):
        texts = [text] if isinstance(text, str) else list(text)
        ids = torch.tensor([self.encode_one(t, max_length) for t in texts], dtype=torch.long)
        return {"input_ids": ids, "length": torch.tensor([max_length] * len(texts))}

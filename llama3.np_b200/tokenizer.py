"""Host tokenizer with the reference's interface and IDENTICAL ids.

`/root/reference/tokenizer.py:5-66` looks every candidate piece up with `list.index` - O(vocabulary)
per lookup, O(V * len^2) per prompt - which becomes the visible cost of a request once the forward
pass runs on a B200.  This class keeps the surface (`Tokenizer(model_path)`, `encode(text, add_bos=True,
add_eos=False)`, `decode(ids)`, `bos_id = 1`, `eos_id = 2`) and the greedy highest-score merge rule, with
a dict from piece to its FIRST index (what `list.index` returns), so ids are the same and a lookup is O(1).
The vocabulary file (`tokenizer.model.np`: JSON with "tokens" and "scores") is the reference's asset and
is read from wherever the caller keeps it; it is not part of this repository.

Kept quirks: characters without a vocabulary entry are dropped; the first best-scoring pair wins ties;
`decode` strips the CHARACTERS of "<s>" and then of "</s>" from both ends (tokenizer.py:65).
"""
from __future__ import annotations

import json
from typing import Dict, List, Sequence


class Tokenizer:
    def __init__(self, model_path: str):
        with open(model_path, encoding="utf-8") as f:
            model = json.load(f)
        self.vocab: List[str] = model["tokens"]
        self.scores: List[float] = model["scores"]
        self.bos_id = 1
        self.eos_id = 2
        self._index: Dict[str, int] = {}
        for i, piece in enumerate(self.vocab):
            self._index.setdefault(piece, i)  # duplicates: the first index, as list.index

    def str_lookup(self, token: str) -> int:
        return self._index.get(token, -1)

    def encode(self, text: str, add_bos: bool = True, add_eos: bool = False) -> List[int]:
        index, vocab, scores = self._index, self.vocab, self.scores
        tokens = [index[ch] for ch in text if ch in index]
        while len(tokens) > 1:
            best_score, best_id, best_at = -1e10, -1, -1
            prev = vocab[tokens[0]]
            for i in range(len(tokens) - 1):
                nxt = vocab[tokens[i + 1]]
                cand = index.get(prev + nxt, -1)
                if cand != -1 and scores[cand] > best_score:  # strict: the first of equal scores wins
                    best_score, best_id, best_at = scores[cand], cand, i
                prev = nxt
            if best_at == -1:
                break
            tokens[best_at : best_at + 2] = [best_id]
        if add_bos:
            tokens.insert(0, self.bos_id)
        if add_eos:
            tokens.append(self.eos_id)
        return tokens

    def decode(self, ids: Sequence[int]) -> str:
        text = "".join(self.vocab[i] for i in ids)
        return text.strip("<s>").strip("</s>")

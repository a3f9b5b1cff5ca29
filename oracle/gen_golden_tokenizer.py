"""Golden vectors for llama3.np_b200/tokenizer.py.  Run in the AUTHORING container:

    python oracle/gen_golden_tokenizer.py

Builds a small SYNTHETIC vocabulary in the reference's file format (the real `tokenizer.model.np` is
the reference's asset and is not copied), runs the unmodified `/root/reference/tokenizer.py` on it and
stores (text, ids, decoded) triples - including duplicate pieces, equal scores, unknown characters and
the `strip` quirk of `decode` - in tests/golden/tokenizer_cases.json.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from tokenizer import Tokenizer as RefTokenizer  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def synthetic_vocab(seed=5):
    rng = random.Random(seed)
    tokens = ["<unk>", "<s>", "</s>"] + [chr(c) for c in range(32, 127) if chr(c) != "q"]  # 'q' has no entry
    scores = [0.0] * len(tokens)
    words = ["the", "he", "th", "in", "er", "an", " t", " a", "re", "on", "at", "en", " the", "ing", " s", "s>", "<s",
             "dream", "dr", "ea", "eam", " I", " have", "ha", "ve", "av", " h", "once", "on", "ce", "up", "pon", " up",
             "time", "ti", "me", "im", " a ", "ll", "ama", "lla", "am", "ma", "ss", "sss", "</", "s>s"]
    for w in words:
        tokens.append(w)  # note: "on" appears twice (duplicate piece -> first index must win)
        scores.append(round(-rng.random() * 10, 3))
    for a, b in [("he", "th"), ("in", "er")]:  # equal scores: the first pair in the text must win
        scores[tokens.index(b)] = scores[tokens.index(a)]
    return {"tokens": tokens, "scores": scores}


if __name__ == "__main__":
    vocab = synthetic_vocab()
    path = os.path.join(GOLD, "tokenizer_vocab_synthetic.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(vocab, f)
    ref = RefTokenizer(path)
    texts = ["I have a dream", "Once upon a time", "the theatre in the inner tent", "quiet question", "", "a",
             "sss<s>the end</s>s", "llama llama ssss", "he thin inert", "  double  spaces  ", "UPPER lower 123 !?"]
    cases = []
    for t in texts:
        for bos, eos in ((True, False), (False, True), (True, True)):
            ids = ref.encode(t, add_bos=bos, add_eos=eos)
            cases.append({"text": t, "add_bos": bos, "add_eos": eos, "ids": ids, "decoded": ref.decode(ids)})
    with open(os.path.join(GOLD, "tokenizer_cases.json"), "w", encoding="utf-8") as f:
        json.dump(cases, f, indent=0)
    print(len(cases), "cases;", cases[0])

"""Question tokenizer and answer vocabulary (host side, integer results must be bit-exact).

Behavioural mirror of ``utils/tokenizer.py`` (Tokenizer, :40-308) and
``data/build_vocab.py`` (AnswerVocabulary, :40-282) of the reference: same class and
method names, same JSON file formats, same ids for every input.  These stay Python on
purpose (BASELINE north_star: "Tokenisation, vocab lookup and answer-index mapping must be
bit-exact"); they produce the int64 id / mask arrays the CUDA text encoder consumes.
"""
from __future__ import annotations

import json
import os
import re
from collections import Counter
from typing import Dict, Iterable, List, Optional, Tuple

PAD_TOKEN, UNK_TOKEN, START_TOKEN, END_TOKEN = "<PAD>", "<UNK>", "<START>", "<END>"
SPECIAL_TOKENS = [PAD_TOKEN, UNK_TOKEN, START_TOKEN, END_TOKEN]
PAD_IDX, UNK_IDX, START_IDX, END_IDX = 0, 1, 2, 3

_NOT_WORD = re.compile(r"[^\w\s']")     # utils/tokenizer.py:116 keeps word chars, whitespace, apostrophes
_SPACES = re.compile(r"\s+")
_ARTICLES = re.compile(r"\b(a|an|the)\b")  # data/build_vocab.py:83
_ANS_PUNCT = re.compile(r"[^\w\s]")


class Tokenizer:
    """Whitespace tokenizer with <PAD>=0, <UNK>=1, <START>=2, <END>=3."""

    def __init__(self, max_length: int = 20, vocab_size: Optional[int] = None):
        self.max_length = max_length
        self.max_vocab_size = vocab_size
        self.word2idx: Dict[str, int] = {t: i for i, t in enumerate(SPECIAL_TOKENS)}
        self.idx2word: Dict[int, str] = {i: t for t, i in self.word2idx.items()}
        self._is_fitted = False

    @property
    def vocab_size(self) -> int:
        return len(self.word2idx)

    @staticmethod
    def preprocess(text: str) -> str:
        return _SPACES.sub(" ", _NOT_WORD.sub(" ", text.lower())).strip()

    def tokenize(self, text: str) -> List[str]:
        return self.preprocess(text).split()

    def build_vocab(self, questions: Iterable[str], min_freq: int = 2) -> None:
        counts: Counter = Counter()
        for q in questions:
            counts.update(self.tokenize(q))
        # first-seen order, then a stable sort by descending count == the reference's ordering
        kept = sorted((w for w, c in counts.items() if c >= min_freq), key=lambda w: -counts[w])
        if self.max_vocab_size is not None:
            kept = kept[: self.max_vocab_size - len(SPECIAL_TOKENS)]
        nxt = len(SPECIAL_TOKENS)
        for w in kept:
            if w not in self.word2idx:
                self.word2idx[w] = nxt
                self.idx2word[nxt] = w
                nxt += 1
        self._is_fitted = True

    def encode(self, text: str, add_special_tokens: bool = True, padding: bool = True,
               truncation: bool = True) -> Tuple[List[int], List[int]]:
        toks = self.tokenize(text)
        if add_special_tokens:
            toks = [START_TOKEN, *toks, END_TOKEN]
        if truncation and len(toks) > self.max_length:
            toks = toks[: self.max_length]
            if add_special_tokens:
                toks[-1] = END_TOKEN
        ids = [self.word2idx.get(t, UNK_IDX) for t in toks]
        mask = [1] * len(ids)
        short = self.max_length - len(ids)
        if padding and short > 0:
            ids += [PAD_IDX] * short
            mask += [0] * short
        return ids, mask

    def batch_encode(self, texts: Iterable[str], add_special_tokens: bool = True
                     ) -> Tuple[List[List[int]], List[List[int]]]:
        pairs = [self.encode(t, add_special_tokens=add_special_tokens) for t in texts]
        return [p[0] for p in pairs], [p[1] for p in pairs]

    def decode(self, token_ids: Iterable[int], skip_special_tokens: bool = True) -> str:
        words = (self.idx2word.get(i, UNK_TOKEN) for i in token_ids)
        return " ".join(w for w in words if not (skip_special_tokens and w in SPECIAL_TOKENS))

    def save(self, filepath: str) -> None:
        with open(filepath, "w", encoding="utf-8") as f:
            json.dump({"word2idx": self.word2idx, "max_length": self.max_length,
                       "max_vocab_size": self.max_vocab_size}, f, indent=2, ensure_ascii=False)

    def load(self, filepath: str) -> None:
        with open(filepath, "r", encoding="utf-8") as f:
            data = json.load(f)
        self.word2idx = data["word2idx"]
        self.idx2word = {int(i): w for w, i in self.word2idx.items()}
        self.max_length = data.get("max_length", self.max_length)
        self.max_vocab_size = data.get("max_vocab_size", self.max_vocab_size)
        self._is_fitted = True


class AnswerVocabulary:
    """Top-N answer <-> class-index mapping; unknown index decodes to "<UNKNOWN>"."""

    def __init__(self, num_answers: int = 1000):
        self.num_answers = num_answers
        self.answer2idx: Dict[str, int] = {}
        self.idx2answer: Dict[int, str] = {}
        self.answer_counts: Dict[str, int] = {}
        self._is_built = False

    @staticmethod
    def preprocess_answer(answer: str) -> str:
        a = _ARTICLES.sub(" ", answer.lower())
        return _SPACES.sub(" ", _ANS_PUNCT.sub("", a)).strip()

    def _install(self, counter: Counter) -> None:
        self.answer_counts = dict(counter)
        for idx, (ans, _) in enumerate(counter.most_common(self.num_answers)):
            self.answer2idx[ans] = idx
            self.idx2answer[idx] = ans
        self._is_built = True

    def build_from_qa_pairs(self, qa_pairs: List[Dict], answer_key: str = "answer",
                            save_path: Optional[str] = None) -> None:
        self._install(Counter(self.preprocess_answer(qa[answer_key]) for qa in qa_pairs))
        if save_path:
            self.save(save_path)

    def build_from_annotations(self, annotations_path: str, save_path: Optional[str] = None) -> None:
        with open(annotations_path, "r", encoding="utf-8") as f:
            data = json.load(f)
        counter: Counter = Counter()
        for ann in data["annotations"]:
            counter[self.preprocess_answer(ann["multiple_choice_answer"])] += 1
            for d in ann.get("answers", []):
                counter[self.preprocess_answer(d["answer"])] += 1
        self._install(counter)
        if save_path:
            self.save(save_path)

    def encode(self, answer: str) -> int:
        return self.answer2idx.get(self.preprocess_answer(answer), -1)

    def decode(self, idx: int) -> str:
        return self.idx2answer.get(idx, "<UNKNOWN>")

    def is_valid_answer(self, answer: str) -> bool:
        return self.preprocess_answer(answer) in self.answer2idx

    def save(self, filepath: str) -> None:
        d = os.path.dirname(filepath)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(filepath, "w", encoding="utf-8") as f:
            json.dump({"num_answers": self.num_answers, "answer2idx": self.answer2idx,
                       "answer_counts": self.answer_counts}, f, indent=2, ensure_ascii=False)

    def load(self, filepath: str) -> None:
        with open(filepath, "r", encoding="utf-8") as f:
            data = json.load(f)
        self.num_answers = data["num_answers"]
        self.answer2idx = data["answer2idx"]
        self.idx2answer = {int(i): a for a, i in self.answer2idx.items()}
        self.answer_counts = data.get("answer_counts", {})
        self._is_built = True

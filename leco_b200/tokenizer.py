"""CLIP byte-level BPE tokenizer for the text-encoder prologue.

The reference tokenizes through transformers' `CLIPTokenizer` (`/root/reference/train_util.py:60-70`:
`tokenizer(prompts, padding="max_length", max_length=tokenizer.model_max_length, truncation=True,
return_tensors="pt").input_ids`).  This is the same published algorithm (OpenAI CLIP `simple_tokenizer.py`; transformers
`tokenization_clip.py`) written against the files a diffusers checkpoint directory ships
(`tokenizer/vocab.json`, `tokenizer/merges.txt`, optional `tokenizer_config.json` / `special_tokens_map.json`):

    whitespace clean-up + NFC + lower-case  ->  split by the CLIP pattern  ->  bytes -> printable unicode alphabet
    ->  BPE merges by rank, last symbol carrying "</w>"  ->  vocabulary ids  ->  <|startoftext|> ids <|endoftext|>,
    truncated to `model_max_length`, padded with the pad token (SD1.x: <|endoftext|>; SD2.x / SDXL tokenizer_2: "!").

`tests/test_text_prologue_cpu.py` pins it id-for-id against transformers' own CLIPTokenizer on a synthetic vocabulary
(no real vocabulary exists offline) and against the committed golden ids made by `tests/golden/make_clip_golden.py`.
"""
from __future__ import annotations

import json
import os
import unicodedata
from functools import lru_cache
from types import SimpleNamespace
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import regex
import torch

BOS, EOS = "<|startoftext|>", "<|endoftext|>"
_PATTERN = regex.compile(
    r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """The reversible byte -> printable character table of GPT-2 / CLIP BPE vocabularies."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def _clean(text: str) -> str:
    """Text normalisation of tokenization_clip.py: drop control characters, every whitespace run -> one space, NFC,
    lower-case.  (CJK ideographs are NOT split apart: that is what OpenAI's tokenizer, transformers' ftfy path and the
    `tokenizers`-backed CLIPTokenizer of the installed transformers do; only the old no-ftfy BasicTokenizer path split them.)"""
    out = []
    for ch in text:
        cp = ord(ch)
        if cp == 0 or cp == 0xFFFD:
            continue
        cat = unicodedata.category(ch)
        if ch in ("\t", "\n", "\r") or cat == "Zs":
            out.append(" ")
        elif cat.startswith("C"):
            continue
        else:
            out.append(ch)
    text = unicodedata.normalize("NFC", "".join(out))
    return " ".join(tok.lower() for tok in text.split())


class ClipTokenizer:
    """The call surface `train_util.text_tokenize` uses, and nothing else of transformers' tokenizer API."""

    def __init__(self, vocab: Dict[str, int], merges: Sequence[Tuple[str, str]], model_max_length: int = 77,
                 pad_token: str = EOS, unk_token: str = EOS):
        self.encoder = dict(vocab)
        self.ranks = {tuple(m): i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.model_max_length = int(model_max_length)
        self.bos_token_id, self.eos_token_id = self.encoder[BOS], self.encoder[EOS]
        self.unk_token_id = self.encoder.get(unk_token, self.eos_token_id)
        if pad_token not in self.encoder:
            raise ValueError(f"pad token {pad_token!r} is not in the vocabulary")
        self.pad_token, self.pad_token_id = pad_token, self.encoder[pad_token]
        self._cache: Dict[str, Tuple[str, ...]] = {BOS: (BOS,), EOS: (EOS,)}
        self._specials = {BOS, EOS, pad_token}
        self._special_split = regex.compile(
            "(" + "|".join(regex.escape(t) for t in sorted(self._specials, key=len, reverse=True)) + ")")

    # ---- loading ---------------------------------------------------------------------------------------------------
    @classmethod
    def from_files(cls, vocab_file: str, merges_file: str, **kw) -> "ClipTokenizer":
        with open(vocab_file, encoding="utf-8") as f:
            vocab = json.load(f)
        with open(merges_file, encoding="utf-8") as f:
            lines = f.read().strip().split("\n")
        if lines and lines[0].startswith("#version"):
            lines = lines[1:]
        merges = [tuple(l.split()) for l in lines if l and len(l.split()) == 2]
        return cls(vocab, merges, **kw)

    @classmethod
    def from_pretrained(cls, directory: str, subfolder: str = "") -> "ClipTokenizer":
        """`directory[/subfolder]` = a diffusers checkpoint's tokenizer folder (model_util.py:110-114 reads the same one)."""
        d = os.path.join(directory, subfolder) if subfolder else directory
        kw = {}

        def token_text(v):
            return v["content"] if isinstance(v, dict) else v
        cfg_path = os.path.join(d, "tokenizer_config.json")
        if os.path.isfile(cfg_path):
            with open(cfg_path, encoding="utf-8") as f:
                cfg = json.load(f)
            if cfg.get("pad_token") is not None:
                kw["pad_token"] = token_text(cfg["pad_token"])
            if cfg.get("unk_token") is not None:
                kw["unk_token"] = token_text(cfg["unk_token"])
            if isinstance(cfg.get("model_max_length"), int) and cfg["model_max_length"] < 10 ** 6:
                kw["model_max_length"] = cfg["model_max_length"]
        sp_path = os.path.join(d, "special_tokens_map.json")
        if os.path.isfile(sp_path):
            with open(sp_path, encoding="utf-8") as f:
                sp = json.load(f)
            if sp.get("pad_token") is not None:
                kw["pad_token"] = token_text(sp["pad_token"])
        return cls.from_files(os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt"), **kw)

    # ---- BPE -------------------------------------------------------------------------------------------------------
    def _bpe(self, token: str) -> Tuple[str, ...]:
        hit = self._cache.get(token)
        if hit is not None:
            return hit
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            best, best_rank = None, None
            for pair in zip(word[:-1], word[1:]):
                rank = self.ranks.get(pair)
                if rank is not None and (best_rank is None or rank < best_rank):
                    best, best_rank = pair, rank
            if best is None:
                break
            first, second = best
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        self._cache[token] = word
        return word

    def tokenize(self, text: str) -> List[str]:
        """Special tokens (bos, eos AND the pad token) are cut out of the raw text first and map to their own ids, as
        transformers does for every registered special token: with the SD2.x / SDXL-2 pad token "!" an exclamation
        mark in a prompt therefore becomes id("!") and never merges ("4k!!" -> 4k</w> ! !), a quirk the trained
        models have always seen."""
        pieces: List[str] = []
        for seg in self._special_split.split(text):
            if seg in self._specials:
                pieces.append(seg)
                continue
            for tok in _PATTERN.findall(_clean(seg)):
                if tok in (BOS, EOS):
                    pieces.append(tok)
                    continue
                mapped = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
                pieces.extend(self._bpe(mapped))
        return pieces

    def encode(self, text: str, max_length: int = None, truncation: bool = True, pad: bool = True) -> List[int]:
        max_length = max_length or self.model_max_length
        ids = [self.encoder.get(p, self.unk_token_id) for p in self.tokenize(text)]
        if truncation and len(ids) > max_length - 2:
            ids = ids[:max_length - 2]
        ids = [self.bos_token_id] + ids + [self.eos_token_id]
        if pad and len(ids) < max_length:
            ids = ids + [self.pad_token_id] * (max_length - len(ids))
        return ids

    def __call__(self, prompts: Union[str, Iterable[str]], padding: str = "max_length", max_length: int = None,
                 truncation: bool = True, return_tensors: str = "pt"):
        if isinstance(prompts, str):
            prompts = [prompts]
        if padding not in ("max_length", False, None):
            raise NotImplementedError("only padding='max_length' (train_util.py:66) is implemented")
        rows = [self.encode(p, max_length, truncation, pad=padding == "max_length") for p in prompts]
        ids = torch.tensor(rows, dtype=torch.long) if return_tensors == "pt" else rows
        return SimpleNamespace(input_ids=ids)

"""Llama-3 tokenizer and chat format for the experimental examples (role of the reference's experimental/models/llama3/tokenizer.py):
a tiktoken BPE over the released ``tokenizer.model`` rank file plus the 256 reserved special tokens, and the header / message / dialog
framing of the instruct models."""
from __future__ import annotations

import os
from typing import Dict, Iterator, List, Literal, Sequence, TypedDict

Role = Literal["system", "user", "assistant"]

# pre-tokenisation regex of the Llama-3 vocabulary (contractions, letter runs, 1-3 digit groups, punctuation runs, whitespace)
_SPLIT_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")
_N_SPECIAL = 256
_NAMED_SPECIALS = ("<|begin_of_text|>", "<|end_of_text|>", "<|reserved_special_token_0|>", "<|reserved_special_token_1|>",
                   "<|reserved_special_token_2|>", "<|reserved_special_token_3|>", "<|start_header_id|>", "<|end_header_id|>",
                   "<|reserved_special_token_4|>", "<|eot_id|>")
_CHUNK_CHARS = 400_000             # tiktoken's regex engine degrades on very long inputs: encode in slices
_MAX_RUN = 25_000                  # ... and on very long runs without (or of) whitespace


class Message(TypedDict):
    role: Role
    content: str


def _special_token_table(first_id: int) -> Dict[str, int]:
    names = list(_NAMED_SPECIALS) + [f"<|reserved_special_token_{i}|>" for i in range(5, _N_SPECIAL - 5)]
    return {name: first_id + offset for offset, name in enumerate(names)}


def _bounded_runs(text: str, limit: int) -> Iterator[str]:
    """Slices of ``text`` none of which contains more than ``limit`` consecutive whitespace (or non-whitespace) characters."""
    begin = run = 0
    kind = None
    for pos, ch in enumerate(text):
        k = ch.isspace()
        run = run + 1 if k == kind else 1
        kind = k
        if run > limit:
            yield text[begin:pos]
            begin, run = pos, 1
    if text:
        yield text[begin:]


class Tokenizer:
    num_reserved_special_tokens = _N_SPECIAL
    pat_str = _SPLIT_PATTERN

    def __init__(self, model_path: str):
        import tiktoken
        from tiktoken.load import load_tiktoken_bpe
        if not os.path.isfile(model_path):
            raise FileNotFoundError(model_path)
        ranks = load_tiktoken_bpe(model_path)
        self.special_tokens = _special_token_table(len(ranks))
        self.model = tiktoken.Encoding(name=os.path.basename(model_path), pat_str=_SPLIT_PATTERN, mergeable_ranks=ranks,
                                       special_tokens=self.special_tokens)
        self.n_words = self.model.n_vocab
        self.bos_id = self.special_tokens["<|begin_of_text|>"]
        self.eos_id = self.special_tokens["<|end_of_text|>"]
        self.pad_id = -1
        self.stop_tokens = {self.eos_id, self.special_tokens["<|eot_id|>"]}

    def encode(self, s: str, *, bos: bool, eos: bool, allowed_special=(), disallowed_special=()) -> List[int]:
        allowed = allowed_special if isinstance(allowed_special, str) else set(allowed_special)
        ids: List[int] = [self.bos_id] if bos else []
        for start in range(0, len(s), _CHUNK_CHARS):
            for piece in _bounded_runs(s[start:start + _CHUNK_CHARS], _MAX_RUN):
                ids += self.model.encode(piece, allowed_special=allowed, disallowed_special=disallowed_special)
        return ids + [self.eos_id] if eos else ids

    def decode(self, t: Sequence[int]) -> str:
        return self.model.decode(list(t))


class ChatFormat:
    """``<|start_header_id|>role<|end_header_id|>\\n\\n content <|eot_id|>`` per message; a dialog prompt ends with an open assistant header."""

    def __init__(self, tokenizer: Tokenizer):
        self.tokenizer = tokenizer
        sp = tokenizer.special_tokens
        self._bot, self._sh, self._eh, self._eot = sp["<|begin_of_text|>"], sp["<|start_header_id|>"], sp["<|end_header_id|>"], sp["<|eot_id|>"]

    def _plain(self, text: str) -> List[int]:
        return self.tokenizer.encode(text, bos=False, eos=False)

    def encode_header(self, message: Message) -> List[int]:
        return [self._sh, *self._plain(message["role"]), self._eh, *self._plain("\n\n")]

    def encode_message(self, message: Message) -> List[int]:
        return [*self.encode_header(message), *self._plain(message["content"].strip()), self._eot]

    def encode_dialog_prompt(self, dialog: Sequence[Message]) -> List[int]:
        ids = [self._bot]
        for m in dialog:
            ids.extend(self.encode_message(m))
        return ids + self.encode_header({"role": "assistant", "content": ""})

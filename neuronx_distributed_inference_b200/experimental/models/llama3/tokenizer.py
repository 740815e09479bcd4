"""Llama-3 tokenizer and chat format for the experimental examples (reference experimental/models/llama3/tokenizer.py:1-222): a
tiktoken BPE over the released ``tokenizer.model`` rank file plus the 256 reserved special tokens, and the header / message / dialog
framing of the instruct models."""
from __future__ import annotations

import os
from typing import Dict, List, Literal, Sequence, TypedDict

Role = Literal["system", "user", "assistant"]


class Message(TypedDict):
    role: Role
    content: str


class Tokenizer:
    num_reserved_special_tokens = 256
    pat_str = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
    _MAX_ENCODE_CHARS = 400_000           # tiktoken's regex engine degrades on very long inputs: encode in slices
    _MAX_NO_WHITESPACE = 25_000

    def __init__(self, model_path: str):
        import tiktoken
        from tiktoken.load import load_tiktoken_bpe
        assert os.path.isfile(model_path), model_path
        ranks = load_tiktoken_bpe(model_path)
        n = len(ranks)
        named = ["<|begin_of_text|>", "<|end_of_text|>", "<|reserved_special_token_0|>", "<|reserved_special_token_1|>",
                 "<|reserved_special_token_2|>", "<|reserved_special_token_3|>", "<|start_header_id|>", "<|end_header_id|>",
                 "<|reserved_special_token_4|>", "<|eot_id|>"]
        special = named + [f"<|reserved_special_token_{i}|>" for i in range(5, self.num_reserved_special_tokens - 5)]
        self.special_tokens: Dict[str, int] = {t: n + i for i, t in enumerate(special)}
        self.model = tiktoken.Encoding(name=os.path.basename(model_path), pat_str=self.pat_str, mergeable_ranks=ranks,
                                       special_tokens=self.special_tokens)
        self.n_words = self.model.n_vocab
        self.bos_id, self.eos_id = self.special_tokens["<|begin_of_text|>"], self.special_tokens["<|end_of_text|>"]
        self.pad_id = -1
        self.stop_tokens = {self.eos_id, self.special_tokens["<|eot_id|>"]}

    def encode(self, s: str, *, bos: bool, eos: bool, allowed_special=(), disallowed_special=()) -> List[int]:
        out: List[int] = []
        for i in range(0, len(s), self._MAX_ENCODE_CHARS):
            for piece in self._split_whitespaces_or_nonwhitespaces(s[i:i + self._MAX_ENCODE_CHARS], self._MAX_NO_WHITESPACE):
                out.extend(self.model.encode(piece, allowed_special=set(allowed_special) if not isinstance(allowed_special, str) else allowed_special,
                                             disallowed_special=disallowed_special))
        return ([self.bos_id] if bos else []) + out + ([self.eos_id] if eos else [])

    def decode(self, t: Sequence[int]) -> str:
        return self.model.decode(list(t))

    @staticmethod
    def _split_whitespaces_or_nonwhitespaces(s: str, max_run: int):
        """Yield slices of ``s`` so that no slice has more than ``max_run`` consecutive whitespace or non-whitespace characters."""
        if not s:
            return
        start, run, is_space = 0, 0, s[0].isspace()
        for i, ch in enumerate(s):
            sp = ch.isspace()
            if sp != is_space:
                run, is_space = 1, sp
            else:
                run += 1
                if run > max_run:
                    yield s[start:i]
                    start, run = i, 1
        yield s[start:]


class ChatFormat:
    def __init__(self, tokenizer: Tokenizer):
        self.tokenizer = tokenizer

    def encode_header(self, message: Message) -> List[int]:
        t = self.tokenizer
        return ([t.special_tokens["<|start_header_id|>"]] + t.encode(message["role"], bos=False, eos=False)
                + [t.special_tokens["<|end_header_id|>"]] + t.encode("\n\n", bos=False, eos=False))

    def encode_message(self, message: Message) -> List[int]:
        t = self.tokenizer
        return self.encode_header(message) + t.encode(message["content"].strip(), bos=False, eos=False) + [t.special_tokens["<|eot_id|>"]]

    def encode_dialog_prompt(self, dialog: Sequence[Message]) -> List[int]:
        """<|begin_of_text|> + every message + the header of the assistant turn the model is asked to complete."""
        out = [self.tokenizer.special_tokens["<|begin_of_text|>"]]
        for m in dialog:
            out += self.encode_message(m)
        return out + self.encode_header({"role": "assistant", "content": ""})

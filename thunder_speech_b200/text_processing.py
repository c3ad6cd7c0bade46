"""Text side of ``src/thunder/text_processing``: the token table with the reference's special-token rules
(``Vocabulary``, vocab.py:18-130), greedy CTC detokenisation (``BatchTextTransformer.decode_prediction``,
transform.py:93-122, what ``BaseCTCModule.predict`` needs) and label ENCODING for the training step
(``encode``, transform.py:65-91: tokenise -> add start/end -> numericalise -> pad), with the reference's three
tokenisers: characters (default), a sentencepiece model, or a custom function.  All host-side string processing.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple, Union

import numpy as np
from torch import Tensor, nn

__all__ = ["Vocabulary", "BatchTextTransformer", "BPETokenizer", "char_tokenizer", "word_tokenizer"]


def char_tokenizer(text: str) -> List[str]:
    """tokenizer.py:114-123."""
    return list(text)


def word_tokenizer(text: str) -> List[str]:
    """tokenizer.py:102-111."""
    return text.split()


class BPETokenizer:
    """sentencepiece pieces of a text (tokenizer.py:23-30)."""

    def __init__(self, model_path: str):
        import sentencepiece

        self.tokenizer = sentencepiece.SentencePieceProcessor()
        self.tokenizer.Load(model_path)

    def __call__(self, text: str) -> List[str]:
        return self.tokenizer.encode_as_pieces(text)


class Vocabulary(nn.Module):
    """Token table: special tokens are appended in the order blank, pad, unknown, start, end, skipping ``None``
    and tokens already present (vocab.py:44-56).  ``blank_idx`` is therefore ``len(tokens)`` by default."""

    def __init__(self, tokens: List[str], blank_token: str = "<blank>", pad_token: Optional[str] = None,
                 unknown_token: Optional[str] = None, start_token: Optional[str] = None,
                 end_token: Optional[str] = None):
        super().__init__()
        self.unknown_token = unknown_token
        self.start_token = start_token
        self.end_token = end_token
        self.blank_token = blank_token
        self.pad_token = pad_token or blank_token
        itos = list(tokens)
        for tok in (blank_token, pad_token, unknown_token, start_token, end_token):
            if tok and tok not in itos:
                itos.append(tok)
        self.itos = itos
        self.stoi = {t: i for i, t in enumerate(itos)}
        self.blank_idx = itos.index(self.blank_token)
        self.pad_idx = itos.index(self.pad_token)

    def numericalize(self, tokens: List[str]) -> Tensor:
        """vocab.py:68-84: without an unknown token, tokens outside the vocabulary are dropped; with one they map to it."""
        import torch

        if self.unknown_token is None:
            tokens = [t for t in tokens if t in self.stoi]
            return torch.tensor([self.stoi[t] for t in tokens], dtype=torch.long)
        unk = self.stoi[self.unknown_token]
        return torch.tensor([self.stoi.get(t, unk) for t in tokens], dtype=torch.long)

    def add_special_tokens(self, tokens: List[str]) -> List[str]:
        """vocab.py:98-112."""
        if self.start_token is not None:
            tokens = [self.start_token] + tokens
        if self.end_token is not None:
            tokens = tokens + [self.end_token]
        return tokens

    def decode_into_text(self, indices) -> List[str]:
        return [self.itos[int(i)] for i in indices]

    def remove_special_tokens(self, text: str) -> str:
        """String replacement in the reference's order (vocab.py:114-130)."""
        text = text.replace(self.blank_token, "")
        text = text.replace(self.pad_token, "")
        if self.start_token is not None:
            text = text.replace(self.start_token, "")
        if self.end_token is not None:
            text = text.replace(self.end_token, "")
        return text


class BatchTextTransformer(nn.Module):
    def __init__(self, tokens: List[str], blank_token: str = "<blank>", pad_token: str = None,
                 unknown_token: str = None, start_token: str = None, end_token: str = None,
                 sentencepiece_model: Optional[str] = None, custom_tokenizer_function=None):
        super().__init__()
        self.vocab = Vocabulary(tokens, blank_token, pad_token, unknown_token, start_token, end_token)
        if custom_tokenizer_function:
            self.tokenizer: Callable[[str], List[str]] = custom_tokenizer_function
        elif sentencepiece_model:
            self.tokenizer = BPETokenizer(str(sentencepiece_model))
        else:
            self.tokenizer = char_tokenizer

    @classmethod
    def from_sentencepiece(cls, output_dir: str) -> "BatchTextTransformer":
        """Vocabulary and tokenizer from a sentencepiece training output folder (``tokenizer.vocab`` + ``tokenizer.model``),
        like the reference's classmethod (text_processing/transform.py:124-150): one piece per line (first tab-separated
        field), sentencepiece's own ``<s>``, ``</s>``, ``<pad>``, ``<unk>`` entries skipped."""
        skip = ("<s>", "</s>", "<pad>", "<unk>")
        with open(f"{output_dir}/tokenizer.vocab", "r", encoding="utf-8") as f:
            pieces = [line.split("\t")[0] for line in f]
        return cls(tokens=[p for p in pieces if p not in skip], sentencepiece_model=f"{output_dir}/tokenizer.model")

    def encode(self, items: List[str], return_length: bool = True, device=None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """List of texts -> padded int64 ``[B, Lmax]`` (pad value ``pad_idx``) and lengths (transform.py:65-91)."""
        import torch
        from torch.nn.utils.rnn import pad_sequence

        encoded = [self.vocab.numericalize(self.vocab.add_special_tokens(self.tokenizer(x))).to(device=device) for x in items]
        batched = pad_sequence(encoded, batch_first=True, padding_value=self.vocab.pad_idx)
        if return_length:
            return batched, torch.tensor([len(it) for it in encoded], dtype=torch.long).to(device=device)
        return batched

    @property
    def num_tokens(self) -> int:
        return len(self.vocab.itos)

    # -- id rows -> strings, vectorised (the per-token Python loop cost ~50 ms per 256 x 300-token batch and capped the
    #    end-to-end rate of predict_stream).  Same string semantics as the reference: join the token strings, then the
    #    word-boundary replacements, then special tokens removed AS SUBSTRINGS (vocab.py:114-130).
    _PUA = 0xF0000   # multi-character tokens (the specials of a character vocabulary) get private-use code points

    def _tables(self):
        t = getattr(self, "_tab", None)
        if t is None or t[0] != len(self.vocab.itos):
            itos = self.vocab.itos
            multi = [(i, tok) for i, tok in enumerate(itos) if len(tok) != 1]
            cp = None
            if len(multi) <= 8 and not any(0xF0000 <= ord(ch) <= 0xF00FF for tok in itos for ch in tok):
                cp = np.array([ord(tok) if len(tok) == 1 else self._PUA + i for i, tok in enumerate(itos)], dtype="<u4")
            obj = np.empty(len(itos), dtype=object)
            obj[:] = itos
            t = self._tab = (len(itos), cp, [(chr(self._PUA + i), tok) for i, tok in multi], obj)
        return t

    def _finish(self, row) -> str:
        _, cp, multi, obj = self._tables()
        row = np.asarray(row)
        if cp is not None:   # character vocabulary: one table lookup + one UTF-32 decode
            out = cp[row].tobytes().decode("utf-32-le")
            for sentinel, tok in multi:
                if sentinel in out:
                    out = out.replace(sentinel, tok)
        else:                # word-piece vocabulary: object-array take + join
            out = "".join(obj[row].tolist())
        out = out.replace("▁", " ")   # sentencepiece word boundary
        out = out.replace("|", " ")   # huggingface word boundary
        return self.vocab.remove_special_tokens(out)

    def decode_collapsed(self, collapsed: Tensor, counts: Tensor) -> List[str]:
        """Detokenise the output of ``thunder_b200::ctc_greedy`` (repeats already collapsed on the GPU):
        ONE device-to-host copy, then the reference's string rules."""
        col = collapsed.cpu().numpy()
        cnt = counts.cpu().numpy()
        return [self._finish(col[b, : cnt[b]]) for b in range(col.shape[0])]

    def decode_prediction(self, predictions: Tensor, remove_repeated: bool = True) -> List[str]:
        """Reference-compatible entry point on argmax ids ``[batch, time]`` (transform.py:93-122)."""
        ids = predictions.detach().cpu().numpy()
        out = []
        for row in ids:
            if remove_repeated and row.size:
                keep = np.ones(row.shape[0], bool)
                keep[1:] = row[1:] != row[:-1]
                row = row[keep]
            out.append(self._finish(row))
        return out

"""Tokenizer and batch sources (reference progen_transformer/data.py).

Tokens: `ord(c) + 1`, 0 is pad / BOS / EOS (data.py:76-88).  Training rows have the reference's output contract
(data.py:64-70): uint16 `(B, seq_len + 1)`, a 0 (BOS) in front, `bytes + 1` truncated to seq_len, zero padded.
The TFRecord/GCS reader itself is outside the hot path (tensorflow is not installable here); `synthetic_iterator`
produces the uniform-random rows BASELINE.json measures on, `iterator_from_sequences` turns strings into rows."""
import numpy as np


def encode_token(token):
    return ord(token) + 1


def decode_token(token):
    if token < 0:
        return ''
    return str(chr(token))


def encode_tokens(tokens):
    return list(map(encode_token, tokens))


def decode_tokens(tokens, offset=1):
    return ''.join(list(map(decode_token, np.asarray(tokens).astype(np.int16) - offset)))


def collate(seqs, seq_len):
    """list of byte strings -> uint16 (B, seq_len + 1) rows exactly like data.py:29-35,64-70"""
    rows = np.zeros((len(seqs), seq_len + 1), np.uint16)
    for i, s in enumerate(seqs):
        t = np.frombuffer(s if isinstance(s, bytes) else s.encode(), dtype=np.uint8).astype(np.uint16)[:seq_len] + 1
        rows[i, 1:1 + len(t)] = t
    return rows


def iterator_from_sequences(seqs, seq_len, batch_size, skip=0, loop=False):
    seqs = list(seqs)[skip:]
    while True:
        for i in range(0, len(seqs), batch_size):
            yield collate(seqs[i:i + batch_size], seq_len)
        if not loop:
            return


def synthetic_iterator(seq_len, batch_size, seed=42, vocab=256):
    """uniform-random [0, vocab) rows of seq_len + 1 tokens (BASELINE.json north_star), endless"""
    rng = np.random.default_rng(seed)
    while True:
        yield rng.integers(0, vocab, (batch_size, seq_len + 1)).astype(np.uint16)

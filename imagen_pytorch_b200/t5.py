"""Text-embedding step in front of the sampling hot path: host-side mirror of the reference's
``imagen_pytorch/t5.py`` -- same entry points (``t5_encode_text``, ``t5_tokenize``, ``t5_encode_tokenized_text``,
``get_encoded_dim``, ``T5_CONFIGS``, ``MAX_LENGTH``, ``DEFAULT_T5_NAME``), same masking (padding embeddings are forced to
zero, t5.py:102-108), so ``Imagen.sample(texts=[...])`` behaves like the reference's.

The T5 encoder itself is a frozen library model (HuggingFace ``transformers``), not part of the denoising loop and not
re-implemented here; what the hot path needs from this module is *fast repeat prompts*:

* ``T5_CONFIGS[name]`` is the same registry the reference uses -- pre-populate ``{'model': ..., 'tokenizer': ...}`` (or call
  ``register_text_encoder``) to use an encoder that is already in memory / on local disk; nothing is downloaded implicitly
  when ``HF_HUB_OFFLINE`` is set or the files are in the local cache;
* ``TextEmbedCache``: an LRU of per-prompt embeddings (unpadded, on the host) keyed by (encoder name, prompt).  ``sample(texts=)``
  looks prompts up first and only runs the encoder on the misses; a batch is re-assembled with the reference's 'longest' padding
  and zeroed pad positions, so cached and uncached batches give identical ``text_embeds`` / ``text_masks``.
"""
from __future__ import annotations

import collections
from typing import List

import torch

MAX_LENGTH = 256                                                               # t5.py:19
DEFAULT_T5_NAME = 'google/t5-v1_1-base'                                        # t5.py:21
T5_CONFIGS = {}                                                                # t5.py:23: name -> {'model', 'tokenizer', 'config'}

# hidden sizes of the encoders the reference documents, used when the HF config cannot be read offline (t5.py:47-58 reads it)
_KNOWN_DIMS = {'t5-small': 512, 't5-base': 768, 't5-large': 1024, 't5-3b': 1024, 't5-11b': 1024,
               'google/t5-v1_1-small': 512, 'google/t5-v1_1-base': 768, 'google/t5-v1_1-large': 1024,
               'google/t5-v1_1-xl': 2048, 'google/t5-v1_1-xxl': 4096}


def register_text_encoder(name, model, tokenizer):
    """Make `name` resolve to an encoder that is already loaded (any module returning `.last_hidden_state`, any tokenizer with
    `batch_encode_plus`); the reference achieves the same by writing into T5_CONFIGS."""
    T5_CONFIGS[name] = dict(model=model, tokenizer=tokenizer, config=getattr(model, 'config', None))


def get_tokenizer(name):
    from transformers import T5Tokenizer
    return T5Tokenizer.from_pretrained(name, model_max_length=MAX_LENGTH)


def get_model(name):
    from transformers import T5EncoderModel
    return T5EncoderModel.from_pretrained(name)


def get_model_and_tokenizer(name):
    entry = T5_CONFIGS.setdefault(name, {})
    if 'model' not in entry:
        entry['model'] = get_model(name)
    if 'tokenizer' not in entry:
        entry['tokenizer'] = get_tokenizer(name)
    return entry['model'], entry['tokenizer']


def get_encoded_dim(name):
    entry = T5_CONFIGS.get(name)
    cfg = None
    if entry is not None:
        cfg = entry.get('config') or getattr(entry.get('model'), 'config', None)
    if cfg is None:
        if name in _KNOWN_DIMS:                                               # no hub round trip for the documented encoders
            return _KNOWN_DIMS[name]
        try:
            from transformers import T5Config
            cfg = T5Config.from_pretrained(name)
            T5_CONFIGS.setdefault(name, {})['config'] = cfg
        except Exception as exc:
            raise ValueError(f'unknown text encoder {name!r} and its config is not available offline: pass text_embed_dim explicitly') from exc
    return cfg.d_model


def t5_tokenize(texts: List[str], name=DEFAULT_T5_NAME):
    model, tokenizer = get_model_and_tokenizer(name)
    if torch.cuda.is_available():
        model = model.cuda()
        T5_CONFIGS[name]['model'] = model
    device = next(model.parameters()).device
    enc = tokenizer.batch_encode_plus(texts, return_tensors='pt', padding='longest', max_length=MAX_LENGTH, truncation=True)
    return enc.input_ids.to(device), enc.attention_mask.to(device)


@torch.no_grad()
def t5_encode_tokenized_text(token_ids, attn_mask=None, pad_id=None, name=DEFAULT_T5_NAME):
    assert attn_mask is not None or pad_id is not None
    model, _ = get_model_and_tokenizer(name)
    if attn_mask is None:
        attn_mask = (token_ids != pad_id).long()
    model.eval()
    hidden = model(input_ids=token_ids, attention_mask=attn_mask).last_hidden_state.detach()
    return hidden.masked_fill(~attn_mask.bool()[..., None], 0.)               # padding embeddings are exactly 0 (t5.py:107)


def t5_encode_text(texts: List[str], name=DEFAULT_T5_NAME, return_attn_mask=False):
    token_ids, attn_mask = t5_tokenize(texts, name=name)
    encoded = t5_encode_tokenized_text(token_ids, attn_mask=attn_mask, name=name)
    return (encoded, attn_mask.bool()) if return_attn_mask else encoded


class TextEmbedCache:
    """LRU of per-prompt text embeddings.  encode(texts, name) == t5_encode_text(texts, name, return_attn_mask=True) for prompts
    that were encoded in ANY earlier batch: T5's encoder output of a prompt does not depend on the other prompts of the batch
    (padding is masked out of its self-attention), only the padded length does -- which is rebuilt here."""

    def __init__(self, capacity=4096, encode_fn=None):
        self.capacity, self.encode_fn = capacity, encode_fn
        self.store = collections.OrderedDict()
        self.hits = self.misses = 0

    def encode(self, texts: List[str], name=DEFAULT_T5_NAME, return_attn_mask=True):
        fn = self.encode_fn or t5_encode_text
        missing = [t for t in dict.fromkeys(texts) if (name, t) not in self.store]
        self.hits += len(texts) - len(missing)
        self.misses += len(missing)
        if missing:
            emb, mask = fn(missing, name=name, return_attn_mask=True)
            emb, mask = emb.float().cpu(), mask.bool().cpu()
            for i, t in enumerate(missing):
                self.store[(name, t)] = emb[i, :int(mask[i].sum())].clone()    # HF pads on the right: the valid tokens are a prefix
        rows = []
        for t in texts:
            self.store.move_to_end((name, t))
            rows.append(self.store[(name, t)])
        while len(self.store) > self.capacity:
            self.store.popitem(last=False)
        longest = max(r.shape[0] for r in rows)
        out = torch.zeros(len(rows), longest, rows[0].shape[-1])
        mask = torch.zeros(len(rows), longest, dtype=torch.bool)
        for i, r in enumerate(rows):
            out[i, :r.shape[0]] = r
            mask[i, :r.shape[0]] = True
        return (out, mask) if return_attn_mask else out

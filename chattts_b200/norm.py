"""Host pre-processing of ``Chat.infer``: text normaliser with the reference's interface and semantics.

Mirrors ``ChatTTS/norm.py:65-253`` (``Normalizer``: ``__call__(text, do_text_normalization, do_homophone_replacement,
lang)``, ``register`` / ``unregister`` / ``destroy``) as plain-Python string work:

* tag-aware application of a registered per-language normaliser (``[uv_break]``-style tags are kept verbatim),
  half-width -> full-width punctuation for Chinese text (norm.py:126-160,163-183);
* invalid-character detection outside tags, punctuation simplification and removal (norm.py:184-201,228-241);
* homophone replacement from ``homophones_map.json`` (norm.py:190-199,220-226) - one ``str.translate`` over a
  code-point table (O(1) per character) instead of the reference's numba scan of the 2 x N table per character;
* language guess: more CJK characters than Latin words => ``zh`` (norm.py:243-253).

No GPU work; this is the "after >= 10x on the model these dominate small requests" row of the scope table.
"""
from __future__ import annotations

import json
import logging
import re
from typing import Callable, Dict, List, Optional, Set, Tuple

_NAMED_TAG = re.compile(r"\[[\w_]+\]")                # tags ignored by the invalid-character count (norm.py:89)
_REJECT = re.compile(r"[^一-鿿A-Za-z，。、,\. ]")  # everything the model cannot read (norm.py:88)
_CJK = re.compile(r"[一-鿿]")
_LATIN_WORD = re.compile(r"\b[A-Za-z]+\b")

# punctuation the model reads as a pause / stop (norm.py:92-125)
_SIMPLIFY = str.maketrans({
    "：": "，", "；": "，", "！": "。", "（": "，", "）": "，", "【": "，", "】": "，", "『": "，", "』": "，",
    "「": "，", "」": "，", "《": "，", "》": "，", "－": "，",
    ":": ",", ";": ",", "!": ".", "(": ",", ")": ",", ">": ",", "<": ",", "-": ",",
})
# half-width ASCII punctuation -> full-width forms for Chinese text; '[', ']' and '_' stay (tags) (norm.py:126-160)
_HALF2FULL = str.maketrans({
    "!": "！", '"': "“", "'": "‘", "#": "＃", "$": "＄", "%": "％", "&": "＆", "(": "（", ")": "）", ",": "，",
    "-": "－", "*": "＊", "+": "＋", ".": "。", "/": "／", ":": "：", ";": "；", "<": "＜", "=": "＝", ">": "＞",
    "?": "？", "@": "＠", "\\": "＼", "^": "＾", "`": "｀", "{": "｛", "|": "｜", "}": "｝", "~": "～",
})


def split_tags(text: str) -> Tuple[List[str], List[str]]:
    """``texts`` between tags and the ``tags`` themselves (norm.py:36-55).  The reference's scanner is reproduced
    character by character, including what it does with malformed input: every '[' closes the running text and
    (re)starts a tag, a ']' ends the tag being collected - or, outside a tag, stays in the text *and* records an
    empty tag - and an unterminated tag is dropped."""
    texts: List[str] = []
    tags: List[str] = []
    cur_text: List[str] = []
    cur_tag: List[str] = []
    for ch in text:
        if ch == "[":
            texts.append("".join(cur_text))
            cur_text = []
            cur_tag = [ch]
        elif cur_tag:
            cur_tag.append(ch)
        else:
            cur_text.append(ch)
        if ch == "]":
            tags.append("".join(cur_tag))
            cur_tag = []
    if cur_text:
        texts.append("".join(cur_text))
    return texts, tags


def combine_tags(texts: List[str], tags: List[str]) -> str:
    """Interleave (norm.py:58-66): text k is followed by tag k while tags last."""
    out = []
    for k, t in enumerate(texts):
        out.append(t)
        if k < len(tags):
            out.append(tags[k])
    return "".join(out)


class Normalizer:
    def __init__(self, map_file_path: Optional[str] = None, logger=logging.getLogger(__name__),
                 homophones: Optional[Dict[str, str]] = None):
        self.logger = logger
        self.normalizers: Dict[str, Callable[[str], str]] = {}
        if homophones is None and map_file_path is not None:
            with open(map_file_path, "r", encoding="utf-8") as f:
                homophones = json.load(f)
        # code point -> code point; only BMP characters can be replaced, like the reference's utf-16 unit table
        self.homophones_map: Dict[int, int] = {ord(k): ord(v) for k, v in (homophones or {}).items()
                                               if len(k) == 1 and len(v) == 1 and ord(k) < 0x10000 and ord(v) < 0x10000}

    # ------------------------------------------------------------------ norm.py:163-201
    def __call__(self, text: str, do_text_normalization: bool = True, do_homophone_replacement: bool = True,
                 lang: Optional[str] = None) -> str:
        if do_text_normalization:
            _lang = self._detect_language(text) if lang is None else lang
            if _lang in self.normalizers:
                texts, tags = split_tags(text)
                texts = [self.normalizers[_lang](t) for t in texts]
                text = combine_tags(texts, tags) if tags else texts[0]
            if _lang == "zh":
                text = text.translate(_HALF2FULL)
        invalid = self._count_invalid_characters(text)
        if invalid:
            self.logger.warning(f"found invalid characters: {invalid}")
            text = text.translate(_SIMPLIFY)
        if do_homophone_replacement and self.homophones_map:
            replaced = [(c, chr(self.homophones_map[ord(c)])) for c in text if ord(c) in self.homophones_map]
            if replaced:
                text = text.translate(self.homophones_map)
                self.logger.info("replace homophones: " + ", ".join(f"{a}->{b}" for a, b in replaced))
        if invalid:
            texts, tags = split_tags(text)
            texts = [_REJECT.sub("", t) for t in texts]
            text = combine_tags(texts, tags) if tags else texts[0]
        return text

    # ------------------------------------------------------------------ norm.py:203-218
    def register(self, name: str, normalizer: Callable[[str], str]) -> bool:
        if name in self.normalizers:
            self.logger.warning(f"name {name} has been registered")
            return False
        try:
            val = normalizer("test string 测试字符串")
            if not isinstance(val, str):
                self.logger.warning("normalizer must have caller type (str) -> str")
                return False
        except Exception as e:  # the reference swallows and reports
            self.logger.warning(e)
            return False
        self.normalizers[name] = normalizer
        return True

    def unregister(self, name: str):
        self.normalizers.pop(name, None)

    def destroy(self):
        self.normalizers.clear()
        self.homophones_map = {}

    # ------------------------------------------------------------------ helpers
    def _count_invalid_characters(self, s: str) -> Set[str]:
        return set(_REJECT.findall(_NAMED_TAG.sub("", s)))

    def _detect_language(self, sentence: str) -> str:
        return "zh" if len(_CJK.findall(sentence)) > len(_LATIN_WORD.findall(sentence)) else "en"

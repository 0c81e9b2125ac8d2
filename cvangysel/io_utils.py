"""Token-stream helpers (reference: cvangysel-common/py/cvangysel/io_utils.py:18,70-148).

`Word` must keep this module path and field order: instances are pickled into the `meta` file by
bin/prepare.py:373-376 and unpickled by bin/train.py:122-126 / bin/query.py:64-68.
"""
import collections
import unicodedata

Word = collections.namedtuple('Word', ['id', 'count'])

_ALLOWED_PUNCTUATION = frozenset('</>')
_latin_cache = {}


def _is_latin_or_sign(character):
    """io_utils.py:70-88: whitespace passes; otherwise the Unicode name must contain LATIN or SIGN."""
    if character.isspace():
        return True
    known = _latin_cache.get(character)
    if known is None:
        name = unicodedata.name(character)
        known = _latin_cache[character] = ('LATIN' in name) or ('SIGN' in name)
    return known


def filter_non_latin_stream(character_stream):
    return (c for c in character_stream if _is_latin_or_sign(c))


def filter_non_alphanumeric_stream(character_stream):
    """io_utils.py:91-98: keep alphanumerics, whitespace and the three markup characters."""
    return (c for c in character_stream if c.isalnum() or c.isspace() or c in _ALLOWED_PUNCTUATION)


def lowercased_stream(iterable):
    return (s.lower() for s in iterable)


def token_stream(unicode_stream, delimiters=(' ', '\t', '\n'), eos_chars=['\n'], eos_token='</s>',
                 ignore_words=[]):
    """io_utils.py:101-138: split on delimiters / end-of-sentence characters, emitting eos_token for each
    end-of-sentence character and once more after a trailing partial sentence."""
    delimiters, eos_chars, ignore_words = set(delimiters), set(eos_chars), set(ignore_words)
    pending = []

    def flush():
        token = ''.join(pending)
        del pending[:]
        if token and token not in ignore_words:
            return token
        return None

    for char in unicode_stream:
        if char in eos_chars or char in delimiters:
            token = flush()
            if token is not None:
                yield token
            if char in eos_chars:
                yield eos_token
        else:
            pending.append(char)
    had_remainder = bool(pending)
    token = flush()
    if token is not None:
        yield token
    if had_remainder and eos_chars:
        yield eos_token


def tokenize_text(text, ignore_words=set()):
    return tuple(token_stream(
        lowercased_stream(filter_non_latin_stream(filter_non_alphanumeric_stream(iter(text)))),
        eos_chars=[], ignore_words=ignore_words))


def translated_token_stream(iterable, words):
    for word in iterable:
        if word in words:
            yield words[word].id

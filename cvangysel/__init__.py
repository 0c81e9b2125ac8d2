"""Minimal stand-in for the `cvangysel` utility package (reference submodule cvangysel-common).

Only what bin/train.py / bin/query.py and the pickled `meta` files need is provided:
argparse validators, logging set-up, `io_utils.Word` (instances are pickled into `meta`,
bin/prepare.py:373-376), the query tokeniser, TREC topic parsing / run writing and the word2vec
binary reader.  Behaviour follows the reference modules cited in each file; heavy third-party
imports of the original package (bs4, nltk, gensim, pyndri) are not needed on this path.
"""
from cvangysel import argparse_utils, embedding_utils, io_utils, logging_utils, sklearn_utils, trec_utils  # noqa: F401

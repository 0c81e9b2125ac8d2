"""argparse type validators (reference: cvangysel-common/py/cvangysel/argparse_utils.py:5-62)."""
import argparse
import os


def _checked(convert, predicate, description):
    def validator(value):
        try:
            converted = convert(value)
        except Exception:
            converted = None
        if converted is None or not predicate(converted):
            raise argparse.ArgumentTypeError('"{0}" is not a valid {1}'.format(value, description))
        return converted
    return validator


# NB: the reference's "positive" int accepts zero (argparse_utils.py:5-13)
positive_int = _checked(int, lambda v: v >= 0, 'positive int')
positive_float = _checked(float, lambda v: v > 0.0, 'positive float')
ratio = _checked(float, lambda v: 0.0 <= v <= 1.0, 'ratio')


def existing_file_path(value):
    path = str(value)
    if not os.path.exists(path):
        raise argparse.ArgumentTypeError('File "{0}" does not exists.'.format(path))
    return path


def nonexisting_file_path(value):
    path = str(value)
    if os.path.exists(path):
        raise argparse.ArgumentTypeError('File "{0}" already exists.'.format(path))
    return path


def bytes(encoding):
    return lambda value: value.encode(encoding)

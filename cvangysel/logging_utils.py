"""Logging set-up used by the CLIs (behaviour of cvangysel-common/py/cvangysel/logging_utils.py:8-63: root logger at
args.loglevel, one line format for every handler, optional <output_path>.log file that must not exist yet)."""
import logging
import os
import subprocess
import sys

LOG_FORMAT = '%(asctime)s [%(threadName)s] [%(name)s] [%(levelname)s]  %(message)s'


def get_formatter():
    return logging.Formatter(LOG_FORMAT)


def _level_from(args):
    name = str(getattr(args, 'loglevel', 'INFO')).upper()
    value = logging.getLevelName(name)          # int for a known name, 'Level X' otherwise
    if not isinstance(value, int):
        raise ValueError('Invalid log level: %s' % name)
    return value


def _attach_log_file(root, formatter, output_path):
    target = '%s.log' % output_path
    if os.path.exists(target):                  # an earlier run wrote here: refuse to mix logs
        logging.error('Model output already exists.')
        raise IOError()
    handler = logging.FileHandler(target)
    handler.setFormatter(formatter)
    root.addHandler(handler)


def configure_logging(args, output_path=None):
    level = _level_from(args)
    logging.basicConfig(level=level)
    root = logging.getLogger()
    root.setLevel(level)
    formatter = get_formatter()
    for existing in list(root.handlers):
        existing.setFormatter(formatter)
    if output_path is not None:
        _attach_log_file(root, formatter, output_path)
    logging.info('Arguments: %s', args)
    logging.info('Git revision: %s', get_git_revision_hash())


def log_module_info(*modules):
    for mod in modules:
        where = getattr(mod, '__path__', None) or getattr(mod, '__file__', '')
        logging.info('%s version: %s (%s)', mod.__name__, getattr(mod, '__version__', 'n/a'), where)


def get_git_revision_hash():
    """HEAD of the checkout the running script lives in, or None outside a git work tree."""
    script_dir = os.path.dirname(os.path.realpath(sys.path[0] or __file__))
    try:
        out = subprocess.run(['git', 'rev-parse', 'HEAD'], cwd=script_dir, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, check=False).stdout
    except OSError:
        return None
    return out.strip() or None

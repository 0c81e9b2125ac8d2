"""Logging configuration (reference: cvangysel-common/py/cvangysel/logging_utils.py:8-63)."""
import logging
import os
import subprocess
import sys

LOG_FORMAT = '%(asctime)s [%(threadName)s] [%(name)s] [%(levelname)s]  %(message)s'


def get_formatter():
    return logging.Formatter(LOG_FORMAT)


def configure_logging(args, output_path=None):
    name = getattr(args, 'loglevel', 'INFO').upper()
    level = getattr(logging, name, None)
    if not isinstance(level, int):
        raise ValueError('Invalid log level: %s' % name)
    logging.basicConfig(level=level)
    root = logging.getLogger()
    root.setLevel(level)
    formatter = get_formatter()
    for handler in root.handlers:
        handler.setFormatter(formatter)
    if output_path is not None:
        log_path = '{0}.log'.format(output_path)
        if os.path.exists(log_path):
            logging.error('Model output already exists.')
            raise IOError()
        file_handler = logging.FileHandler(log_path)
        file_handler.setFormatter(formatter)
        root.addHandler(file_handler)
    logging.info('Arguments: %s', args)
    logging.info('Git revision: %s', get_git_revision_hash())


def log_module_info(*modules):
    for module in modules:
        logging.info('%s version: %s (%s)', module.__name__,
                     getattr(module, '__version__', 'n/a'), getattr(module, '__path__', getattr(module, '__file__', '')))


def get_git_revision_hash():
    try:
        here = os.path.dirname(os.path.realpath(sys.path[0] or __file__))
        return subprocess.Popen(['git', 'rev-parse', 'HEAD'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                cwd=here).communicate()[0].strip()
    except Exception:
        return None

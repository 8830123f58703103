"""diffusers.utils.logging verbosity helpers (ddpm_train.py:279-281)."""
import logging as _logging

_root = _logging.getLogger("diffusers")


def get_logger(name=None):
    return _logging.getLogger(name or "diffusers")


def set_verbosity(level):
    _root.setLevel(level)


def set_verbosity_info():
    set_verbosity(_logging.INFO)


def set_verbosity_warning():
    set_verbosity(_logging.WARNING)


def set_verbosity_error():
    set_verbosity(_logging.ERROR)


def set_verbosity_debug():
    set_verbosity(_logging.DEBUG)

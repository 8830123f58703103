"""accelerate.logging.get_logger (ddpm_train.py:13,26,271,318): a logging adapter whose calls accept `main_process_only=`."""
import logging
import os


class _Adapter(logging.LoggerAdapter):
    def log(self, level, msg, *args, main_process_only=True, **kwargs):
        if main_process_only and int(os.environ.get("RANK", "0")) != 0:
            return
        kwargs.pop("in_order", None)
        if self.isEnabledFor(level):
            msg, kwargs = self.process(msg, kwargs)
            self.logger.log(level, msg, *args, **kwargs)


def get_logger(name: str, log_level: str = None):
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level.upper())
        logger.root.setLevel(log_level.upper())
    return _Adapter(logger, {})

"""fatal()/warning() with the reference's behaviour (opengate/exception.py:46-61): fatal raises
a plain Exception after logging, warning only logs."""

import logging

logger = logging.getLogger("opengate_b200")


class GateImplementationError(Exception):
    pass


def fatal(msg):
    logger.critical(msg)
    raise Exception(msg)


def warning(msg):
    logger.warning(msg)

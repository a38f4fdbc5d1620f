import logging
import sys

_FMT = "%(asctime)s %(levelname)s %(name)s: %(message)s"


def get_logger(name="ever", level=logging.INFO, **kw):
    lg = logging.getLogger(name)
    if not lg.handlers:
        h = logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter(_FMT))
        lg.addHandler(h)
        lg.setLevel(level)
        lg.propagate = False
    return lg


def info(msg, *a):
    get_logger().info(msg, *a)

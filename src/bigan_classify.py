"""`PYTHONPATH=src` compatibility shim (reference runtests:3): the implementation lives in
cellcomm_b200.bigan_classify."""
from cellcomm_b200.bigan_classify import *  # noqa: F401,F403
from cellcomm_b200.bigan_classify import __dict__ as _d

globals().update({k: v for k, v in _d.items() if not k.startswith('__')})
